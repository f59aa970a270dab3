#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
