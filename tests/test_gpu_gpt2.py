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


@pytest.mark.parametrize("kernels", ["fast", "generic"])
def test_gpt2_logits_track_cpu_step_by_step(model, kernels):
    """End-to-end parity of the whole token graph, QUANTIFIED: ggml-cpu generates N_STEPS greedy tokens and dumps its logits; the B200
    backend is then teacher-forced along the SAME trajectory (so step i sees the identical context on both sides) and its logits are
    compared step by step.  Asserted, for the default (superblock / fused) kernels and for the generic kernels:
      * NMSE(logits_b200, logits_cpu) <= 1e-6 at every step (the reference's own per-op gates are 1e-7 .. 5e-4);
      * the greedy token is IDENTICAL whenever the CPU's top-2 margin exceeds the measured logit noise of that step (4 x the largest
        absolute deviation) -- i.e. any token difference is a genuine tie inside the f32 summation-order noise, not a bug;
      * free-running (unforced) generation matches the CPU up to the first such tie."""
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
    ties, worst_nmse, worst_ratio = [], 0.0, 0.0
    for i in range(N_STEPS):
        c, g_ = cpu_logits[i], gpu_logits[i]
        nm = O.nmse(g_, c)
        worst_nmse = max(worst_nmse, nm)
        assert nm <= 1e-6, (kernels, i, nm)
        noise = float(np.abs(g_ - c).max())
        top2 = np.sort(c)[-2:]
        margin = float(top2[1] - top2[0])
        worst_ratio = max(worst_ratio, noise / max(margin, 1e-30))
        if int(g_.argmax()) != int(c.argmax()):
            ties.append((i, margin, noise))
            assert margin <= 4 * noise, f"{kernels}: step {i}: argmax differs although the CPU margin {margin:.3e} exceeds 4 x the logit noise {noise:.3e}"
    # free-running generation: identical to the CPU until the first tie (if any)
    _, free_logits, _ = run_dump("gpt-2-backend-b200-dump", model, N_STEPS, TMP / f"b200_{kernels}_free.logits", extra=("-ngl", "12"), env_extra=env)
    ftoks = [int(v) for v in free_logits.argmax(1)]
    first_tie = ties[0][0] if ties else N_STEPS
    assert ftoks[:first_tie] == ctoks[:first_tie], f"\ncpu: {ctoks}\ngpu: {ftoks}\nties: {ties}"
    print(f"gpt-2 117M q4_0 [{kernels} kernels]: {N_STEPS} teacher-forced steps, worst logits NMSE {worst_nmse:.2e}, worst noise/margin {worst_ratio:.2e}, "
          f"argmax differs at {len(ties)} steps (all inside the noise): {ties[:4]}; free-running prefix identical for {first_tie} tokens")


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
