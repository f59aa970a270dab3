"""GPU: the reference's examples/gpt-2 programs, compiled UNMODIFIED, running a synthetic GPT-2 117M (Q4_0) on the
B200 backend (BASELINE.json configs[3]).  `gpt-2-backend-b200` = main-backend.cpp built with -DGGML_USE_CUDA and linked
against libggml-b200.so (which exports the ggml-cuda.h ABI), so every node of the token graph runs on the device
(no scheduler, no CPU fallback).  Greedy decoding (--top_k 1) must produce the same tokens as the CPU backend."""
import re
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
TMP = Path("/tmp/ggml_b200_gpt2_v2")


@pytest.fixture(scope="module")
def model():
    for exe in ("gpt-2-backend", "gpt-2-backend-b200", "gpt-2-sched-b200", "gpt-2-quantize"):
        if not (O.REF_DIR / exe).exists():
            pytest.fail(f"oracle/_ref/{exe} missing (make -C oracle ref b200bins in the build container)")
    TMP.mkdir(exist_ok=True)
    f16, q40 = TMP / "gpt2_f16.bin", TMP / "gpt2_q4_0.bin"
    if not q40.exists():
        subprocess.run([sys.executable, str(ROOT / "scripts" / "make_gpt2_synth.py"), str(f16)], check=True)
        subprocess.run([str(O.REF_DIR / "gpt-2-quantize"), str(f16), str(q40), "2"], check=True, env=O.ref_env(), capture_output=True)
    return q40


def run(exe, model, n=24, extra=(), env_extra=None):
    cmd = [str(O.REF_DIR / exe), "-m", str(model), "-s", "1234", "-n", str(n), "-t", "8", "--ignore-eos", "--top_k", "1", "-p", "a b c", *extra]
    env = O.ref_env()
    if env_extra:
        env.update(env_extra)
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-2000:]
    text = [l for l in p.stdout.splitlines() if l.startswith("a b c")]
    ms = re.search(r"predict time =\s*([\d.]+) ms /\s*([\d.]+) ms per token", out)
    return text[0] if text else "", float(ms.group(2)) if ms else None, out


def run_dump(exe, model, n, dump, force=None, extra=(), env_extra=None):
    """the "-dump" variants (oracle/gpt2_logits_hook.cpp): main-backend.cpp unmodified, logits of every sampling step written to `dump`,
    optionally teacher-forced with the token ids in `force`"""
    env_extra = dict(env_extra or {})
    env_extra["GPT2_LOGITS_DUMP"] = str(dump)
    if force is not None:
        env_extra["GPT2_FORCE_TOKENS"] = str(force)
    text, ms, out = run(exe, model, n=n, extra=extra, env_extra=env_extra)
    import numpy as np
    logits = np.fromfile(dump, dtype=np.float32).reshape(-1, 50257)
    return text, logits, out


N_STEPS = 48


def run_compare(model, mode, n_prompt=5, dev="B2000"):
    """oracle/_ref/gpt2-compare: the example's own graph (main-backend.cpp included unmodified) evaluated node by node on ggml-cpu and on `dev` with
    the reference's ggml_backend_compare_graph_backend; returns {phase: (nodes over 1e-9, worst NMSE, first node over, its op)} and the per-node lines"""
    import ggml_b200
    exe = O.REF_DIR / "gpt2-compare"
    if not exe.exists():
        pytest.fail("oracle/_ref/gpt2-compare missing (make -C oracle b200bins in the build container)")
    env = O.ref_env()
    env["GGML_BACKEND_PATH"] = str(ggml_b200.BACKEND_SO)
    cmd = [str(exe), str(model), dev, str(n_prompt)] + (["sync"] if mode == "sync" else [])
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    out = {}
    for l in p.stdout.splitlines():
        f = l.split()
        if f and f[0] == "summary":
            out[f[1]] = (int(f[4]), float(f[6]), int(f[8]), f[9])
    nodes = [l.split() for l in p.stdout.splitlines() if l.startswith("node ")]
    return out, nodes


def test_gpt2_graph_every_node_matches_cpu_on_identical_inputs(model):
    """Per-op parity on the REAL token graph (prompt batch of 5 and one decode step, 487 nodes each): every node evaluated on the B200 backend
    with exactly the inputs the CPU backend had (the device copy of each result is replaced by the CPU's after the comparison) must agree
    with ggml-cpu to NMSE <= 1e-9 -- quantized mat-muls, float mat-muls over the KV cache, norms, soft-max, GELU, copies, adds."""
    out, nodes = run_compare(model, "sync")
    assert set(out) == {"prompt", "decode"}, out
    for phase, (n_over, worst, first, op) in out.items():
        assert n_over == 0 and worst <= 1e-9, (phase, n_over, worst, first, op)
    assert len(nodes) > 500
    print(f"gpt-2 graph, identical inputs per node: worst NMSE prompt {out['prompt'][1]:.2e}, decode {out['decode'][1]:.2e} over {len(nodes)} compared nodes")


def test_gpt2_graph_free_running_deviation_starts_at_a_discontinuity(model):
    """Free-running (the device consumes its own intermediate results, as in a real run) the deviation from ggml-cpu is NOT 1e-7: the
    reference algorithm contains discontinuous steps -- GELU through an f16 -> f16 table (ggml-cpu.c:1355 ff.) and the int8 quantization of
    the activations in front of every quantized mat-mul (ggml-cpu.c:7490-7509) -- which turn a 1e-7 difference in f32 summation order into an
    occasional jump of one f16 ulp (5e-4) or one int8 code (1e-2) on single elements, and those cascade from layer to layer to ~1e-4 NMSE on
    the logits.  Asserted: wherever a deviation above 1e-9 first appears it is at such a node (GELU, or a MUL_MAT fed by one), everything
    before it is below 1e-9, and the accumulated deviation stays bounded (<= 5e-3)."""
    out, nodes = run_compare(model, "free")
    for phase, (n_over, worst, first, op) in out.items():
        assert worst <= 5e-3, (phase, worst)
        if n_over:
            assert op in ("GELU", "MUL_MAT"), (phase, first, op)
    print("gpt-2 graph, free-running: " + ", ".join(f"{ph}: {v[0]} nodes over 1e-9, worst {v[1]:.2e}, first at node {v[2]} ({v[3]})" for ph, v in out.items()))


@pytest.mark.parametrize("kernels", ["fast", "generic"])
def test_gpt2_logits_track_cpu_step_by_step(model, kernels):
    """End-to-end, QUANTIFIED: ggml-cpu generates N_STEPS greedy tokens and dumps its logits; the B200 backend is teacher-forced along the SAME
    trajectory (step i sees the identical context on both sides) and its logits are compared step by step.  Asserted, for the default
    (superblock / fused) kernels and for the generic kernels:
      * NMSE(logits_b200, logits_cpu) <= 5e-3 at every step -- the bound of the cascade through the reference's discontinuous steps (see
        test_gpt2_graph_free_running_deviation_starts_at_a_discontinuity; per-op parity on identical inputs is <= 1e-9);
      * the greedy token is IDENTICAL whenever the CPU's top-2 margin exceeds 6 x the RMS logit deviation of that step, i.e. every token
        difference is a near-tie inside the measured noise, not a wrong computation;
      * free-running (unforced) generation matches the CPU up to the first such near-tie."""
    import numpy as np
    for exe in ("gpt-2-backend-dump", "gpt-2-backend-b200-dump"):
        if not (O.REF_DIR / exe).exists():
            pytest.fail(f"oracle/_ref/{exe} missing (make -C oracle b200bins in the build container)")
    env = {"GGML_B200_FORCE_GENERIC": "1"} if kernels == "generic" else None
    cpu_text, cpu_logits, _ = run_dump("gpt-2-backend-dump", model, N_STEPS, TMP / "cpu.logits")
    assert cpu_logits.shape[0] == N_STEPS, cpu_logits.shape
    ctoks = [int(v) for v in cpu_logits.argmax(1)]                              # greedy (--top_k 1): the sampled token is the arg-max of the dumped logits
    printed = [int(t) for t in re.findall(r"<(\d+)>", cpu_text)]                # ids outside the printable range are echoed as <id>
    assert all(t in ctoks for t in printed), (printed, ctoks)
    np.array(ctoks, dtype=np.int32).tofile(TMP / "force.bin")
    _, gpu_logits, out = run_dump("gpt-2-backend-b200-dump", model, N_STEPS, TMP / f"b200_{kernels}.logits", force=TMP / "force.bin", extra=("-ngl", "12"), env_extra=env)
    assert "using CUDA backend" in out, out[-1500:]
    assert gpu_logits.shape == cpu_logits.shape
    ties, worst_nmse, n_same = [], 0.0, 0
    for i in range(N_STEPS):
        c, g_ = cpu_logits[i].astype(np.float64), gpu_logits[i].astype(np.float64)
        nm = O.nmse(gpu_logits[i], cpu_logits[i])
        worst_nmse = max(worst_nmse, nm)
        assert nm <= 5e-3, (kernels, i, nm)
        rms = float(np.sqrt(np.mean((g_ - c) ** 2)))
        top2 = np.sort(c)[-2:]
        margin = float(top2[1] - top2[0])
        if int(g_.argmax()) != int(c.argmax()):
            ties.append((i, margin, rms))
            assert margin <= 6 * rms, f"{kernels}: step {i}: argmax differs although the CPU margin {margin:.3e} exceeds 6 x the RMS logit deviation {rms:.3e}"
        else:
            n_same += 1
    # free-running generation: identical to the CPU until the first near-tie (if any)
    _, free_logits, _ = run_dump("gpt-2-backend-b200-dump", model, N_STEPS, TMP / f"b200_{kernels}_free.logits", extra=("-ngl", "12"), env_extra=env)
    ftoks = [int(v) for v in free_logits.argmax(1)]
    first_tie = ties[0][0] if ties else N_STEPS
    assert ftoks[:first_tie] == ctoks[:first_tie], f"\ncpu: {ctoks}\ngpu: {ftoks}\nties: {ties}"
    assert n_same >= N_STEPS // 2, (n_same, ties)
    print(f"gpt-2 117M q4_0 [{kernels} kernels]: {N_STEPS} teacher-forced steps, worst logits NMSE {worst_nmse:.2e}, same greedy token at {n_same}/{N_STEPS} steps, "
          f"differences only at near-ties (margin, rms): {[(i, round(m, 4), round(r, 4)) for i, m, r in ties[:5]]}; free-running prefix identical for {first_tie} tokens")


def test_gpt2_graph_fusions_are_bit_exact(model):
    """the graph-level fusions (mat-vec + bias + GELU / + residual, norm + gain + bias, scale + mask + soft-max, K/V cache copies) and the
    launch mode (CUDA-graph capture vs direct PDL-chained launches) must not change a single logit bit: teacher-forced along one trajectory"""
    import numpy as np
    toks = np.arange(40, 40 + 16, dtype=np.int32)
    toks.tofile(TMP / "force16.bin")
    outs = {}
    for name, env in (("default", {}), ("nofusion", {"GGML_B200_DISABLE_FUSION": "1"}), ("nographs", {"GGML_B200_DISABLE_GRAPHS": "1"}),
                      ("always_capture", {"GGML_B200_GRAPH_MAX_UPDATES": "1000000"})):
        _, logits, _ = run_dump("gpt-2-backend-b200-dump", model, 16, TMP / f"b200_{name}.logits", force=TMP / "force16.bin", extra=("-ngl", "12"), env_extra=env)
        outs[name] = logits
    for name in ("nofusion", "nographs", "always_capture"):
        assert np.array_equal(outs["default"].view(np.uint32), outs[name].view(np.uint32)), name


def test_gpt2_sched_full_offload(model):
    cpu_text, _, _ = run("gpt-2-backend", model)
    gpu_text, gpu_ms, out = run("gpt-2-sched-b200", model, extra=("-ngl", "99"))
    ctoks, gtoks = re.findall(r"<(\d+)>", cpu_text), re.findall(r"<(\d+)>", gpu_text)
    assert len(ctoks) >= 8 and ctoks[:4] == gtoks[:4], f"\ncpu: {cpu_text}\ngpu: {gpu_text}\n{out[-1500:]}"
