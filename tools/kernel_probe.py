#!/usr/bin/env python
"""Time single kernel forms of libdpmsolver_b200.so in isolation (CUDA events, rotating buffer sets
larger than L2) -- used for tuning sweeps and as the short command wrapped by ncu.

    python tools/kernel_probe.py --form ms3 --dtype bf16 --n-model 0 [--variant 0 --threads 256 --ctas 8]
    python tools/kernel_probe.py --sweep            # table over variants / threads / CTAs per SM
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dpm_solver_b200 import ops  # noqa: E402
from dpm_solver_b200.ops import StepArgs  # noqa: E402

FORMS = {"none": 0, "lin1": 1, "lin2": 2, "lin3": 3, "diff2": 4, "ms3": 5, "ss3t": 6}
DT = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}


def build(form, n_model, sdt, mdt, n, sets=3, m_out=True, thr_ps=0):
    out = []
    f = FORMS[form]
    for _ in range(sets):
        a = StepArgs(form=f, n_model=n_model, predict_x0=n_model > 0, guidance=7.5, alpha_e=0.83, sigma_e=0.55,
                     a=0.95, c0=-0.1, c1=0.05, c2=-0.01, w0=1.02, w1=0.98, w2=0.51, w3=0.5, w4=0.33, want_m_out=m_out,
                     state_dtype=sdt)
        mk = lambda dt: torch.randn(n, device="cuda", dtype=dt)
        if f != 0:
            a.x = mk(sdt)
            a.out = torch.empty(n, device="cuda", dtype=sdt)
        if n_model == 0:
            a.m0 = mk(sdt)
        else:
            a.e_cond = mk(mdt)
            if n_model == 2:
                a.e_uncond = mk(mdt)
            a.xe = a.x if f != 0 else mk(sdt)
            if m_out or f == 0:
                a.m_out = torch.empty(n, device="cuda", dtype=sdt)
        if f in (2, 3, 4, 5, 6):
            a.m1 = mk(sdt)
        if f in (3, 5, 6):
            a.m2 = mk(sdt)
        if thr_ps:
            a.per_sample = thr_ps
            a.thr = torch.rand(n // thr_ps, device="cuda") + 0.5
        out.append(a)
    return out


def algo_bytes(a):
    n = a.reference_tensor().numel()
    seen, tot = set(), 0
    for t in (a.x, a.xe, a.m0, a.m1, a.m2, a.e_cond, a.e_uncond, a.m_out, a.out, a.out2):
        if t is not None and t.data_ptr() not in seen:
            seen.add(t.data_ptr())
            tot += n * t.element_size()
    return tot


def time_form(be, sets, reps=30, warm=5):
    for i in range(warm):
        be.step(sets[i % len(sets)])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (s, e) in enumerate(ev):
        s.record()
        be.step(sets[i % len(sets)])
        e.record()
    torch.cuda.synchronize()
    us = sorted(s.elapsed_time(e) * 1e3 for s, e in ev)
    return us[len(us) // 2], us[0]


def time_chain(be, form, sdt, n, reps=40, fresh=False):
    """Emulate the sampling loop's data flow: x_{i+1} = out_i, m1_{i+1} = m_out_i, eps from rotating
    banks; `fresh` allocates the outputs per step like the solver does (else a fixed ring of 4)."""
    f = FORMS[form]
    mk = lambda: torch.randn(n, device="cuda", dtype=sdt)
    banks = [mk() for _ in range(3)]
    x, m1, m2 = mk(), mk(), mk()
    ring = [torch.empty(n, device="cuda", dtype=sdt) for _ in range(8)]
    ev = []
    for i in range(reps + 5):
        a = StepArgs(form=f, n_model=1, predict_x0=True, alpha_e=0.83, sigma_e=0.55, a=0.95, c0=-0.1, c1=0.05,
                     c2=-0.01, w0=1.02, w1=0.98, w2=0.51, w3=0.5, want_m_out=True, state_dtype=sdt,
                     x=x, xe=x, e_cond=banks[i % 3], m1=m1, m2=m2 if f == 5 else None)
        if not fresh:
            a.out, a.m_out = ring[(2 * i) % 8], ring[(2 * i + 1) % 8]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        m_new, out = be.step(a)
        e.record()
        ev.append((s, e))
        x, m2, m1 = out, m1, m_new
    torch.cuda.synchronize()
    us = sorted(s.elapsed_time(e) * 1e3 for s, e in ev[5:])
    return us[len(us) // 2], us[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--form", default="ms3", choices=sorted(FORMS))
    ap.add_argument("--dtype", default="bf16", choices=sorted(DT))
    ap.add_argument("--model-dtype", default=None)
    ap.add_argument("--n-model", type=int, default=0)
    ap.add_argument("--no-m-out", action="store_true")
    ap.add_argument("--shape", default="4096,4,64,64")
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--thr", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--chain", action="store_true")
    ap.add_argument("--peak", type=float, default=6574.5)
    a = ap.parse_args()
    shape = [int(v) for v in a.shape.split(",")]
    n = 1
    for v in shape:
        n *= v
    be = ops.CudaBackend()
    if a.chain:
        for form in ("diff2", "ms3"):
            for dt in ("bf16", "f32"):
                nb = 5 if form == "diff2" else 6
                b = nb * n * (2 if dt == "bf16" else 4)
                for variant, threads, ctas in ((0, 0, 0), (1, 256, 2), (1, 128, 4), (1, 512, 1)):
                    for fresh in (False, True):
                        be.set_tuning(variant, threads, ctas)
                        med, mn = time_chain(be, form, DT[dt], n, fresh=fresh)
                        print(json.dumps(dict(mode="chain", form=form, dtype=dt, variant=variant, threads=threads, ctas=ctas,
                                              fresh=fresh, median_us=round(med, 1), min_us=round(mn, 1), gbs=round(b / med / 1e3),
                                              frac=round(b / med / 1e3 / a.peak, 3))), flush=True)
        be.set_tuning(2, 0, 0)
        return
    if not a.sweep:
        sdt = DT[a.dtype]
        mdt = DT[a.model_dtype] if a.model_dtype else sdt
        sets = build(a.form, a.n_model, sdt, mdt, n, m_out=not a.no_m_out, thr_ps=(n // shape[0]) if a.thr else 0)
        be.set_tuning(a.variant, a.threads, a.ctas)
        med, mn = time_form(be, sets, a.reps)
        b = algo_bytes(sets[0])
        print(json.dumps({"form": a.form, "dtype": a.dtype, "n_model": a.n_model, "variant": a.variant, "threads": a.threads,
                          "ctas": a.ctas, "bytes": b, "median_us": med, "min_us": mn, "gbs": b / med / 1e3,
                          "frac": b / med / 1e3 / a.peak}))
        return
    rows = []
    cases = [("ms3", 0, "bf16", True), ("ms3", 0, "f32", True), ("diff2", 1, "bf16", True), ("diff2", 2, "bf16", True),
             ("ms3", 1, "f32", True), ("ms3", 2, "bf16", True), ("lin1", 1, "bf16", True), ("diff2", 1, "f32", True),
             ("none", 2, "bf16", True)]
    for form, nm, dt, mo in cases:
        sets = build(form, nm, DT[dt], DT[dt], n, m_out=mo)
        b = algo_bytes(sets[0])
        for variant, threads, ctas in itertools.chain(
                itertools.product([0], [128, 256, 512], [2, 4, 8]),
                itertools.product([1], [128, 256, 512], [1, 2, 3, 4])):
            if threads * ctas > 2048:
                continue
            be.set_tuning(variant, threads, ctas)
            med, mn = time_form(be, sets, 20, 3)
            rows.append(dict(form=form, n_model=nm, dtype=dt, variant=variant, threads=threads, ctas=ctas, median_us=round(med, 1),
                             gbs=round(b / med / 1e3, 1), frac=round(b / med / 1e3 / a.peak, 3)))
            print(json.dumps(rows[-1]), flush=True)
        del sets
        torch.cuda.empty_cache()
    be.set_tuning(2, 0, 0)


if __name__ == "__main__":
    main()
