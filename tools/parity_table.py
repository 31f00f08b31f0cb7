#!/usr/bin/env python
"""Per-variable parity table: ours vs the reference build vs the double-precision trajectory.

For every config x iteration count x variable it prints three norm-wise relative errors
    ours-ref  = ||ours - ref|| / ||ref||        agreement with the reference's own CUDA/cuBLAS build (oracle/_ref/ref_driver)
    ours-f64  = ||ours - f64|| / ||f64||        accuracy: distance from the same algorithm run in double (oracle, -DORC_F64)
    ref-f64   = ||ref  - f64|| / ||f64||        the reference's own rounding uncertainty (the floor any fp32 code lives on)
and the verdict of the two gates of tests/refcompare.py (accuracy: ours-f64 <= ACC_FACTOR ref-f64 + ACC_ABS per variable and a median ratio <= ACC_MEDIAN; agreement: ours-ref <= 1e-4
or floor-limited).  All three run on identical JSON inputs at EQUAL iteration counts (cold start each).

  python tools/parity_table.py --configs toy,C1,C1r6,C1r30,C2,C3 --iters 1,10,100,500 --out profiles/r02_parity_table

Test infrastructure: needs a GPU, oracle/_ref/ref_driver and the oracle port; nothing here is on the product path.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refcompare import ACC_ABS, ACC_FACTOR, ACC_MEDIAN, RTOL, median_ratio, rel_err  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
# (label, our buffer, reference dump, oracle name)
VARS = [("U", "VEC_U", "U", "U"), ("X", "VEC_X", "X", "X"),
        ("y_xi", "VEC_UPDATE_XI", "updateXi", "update_xi"), ("y_psi", "VEC_UPDATE_PSI", "updatePsi", "update_psi"),
        ("yprev_xi", "VEC_XI", "xi", "xi"), ("yprev_psi", "VEC_PSI", "psi", "psi"),
        ("z_xi", "VEC_DUAL_XI", "dualXi", "dual_xi"), ("z_psi", "VEC_DUAL_PSI", "dualPsi", "dual_psi"),
        ("Hx_xi", "VEC_PRIMAL_XI", "primalXi", "primal_xi"), ("Hx_psi", "VEC_PRIMAL_PSI", "primalPsi", "primal_psi"),
        ("w_xi", "VEC_ACCEL_XI", "accelXi", "accel_xi"), ("w_psi", "VEC_ACCEL_PSI", "accelPsi", "accel_psi")]


def load_problem(case, iters):
    from rapidnet_b200.datagen import named_problem
    if case == "toy":
        from rapidnet_b200.problem import problem_from_npz_dict
        z = dict(np.load(os.path.join(ROOT, "tests", "golden", "toy.npz"), allow_pickle=False))
        prob = copy.deepcopy(problem_from_npz_dict(z))
        prob.config.max_iter = iters
        return prob, 1
    return named_problem(case, max_iter=iters), 0


def run_reference(prob, slot):
    from rapidnet_b200.problem import write_problem
    tmp = tempfile.mkdtemp(prefix="rn_par_")
    cfg = write_problem(prob, tmp)
    dump = os.path.join(tmp, "dump")
    os.makedirs(dump)
    t0 = time.perf_counter()
    out = subprocess.run([REF, cfg, str(slot), "0", "1", dump], capture_output=True, text=True, timeout=3000)
    if out.returncode != 0:
        raise RuntimeError(out.stderr[-1000:] + out.stdout[-1000:])
    res = {f[:-4]: np.fromfile(os.path.join(dump, f), dtype=np.float32) for f in os.listdir(dump)}
    subprocess.run(["rm", "-rf", tmp])
    return res, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="toy,C1,C1r6,C1r30,C2,C3")
    ap.add_argument("--iters", default="1,10,100,500")
    ap.add_argument("--factors", default="full", help="comma list of full,df,shared (ours)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity_table"))
    args = ap.parse_args()
    from oracle.oracle import Oracle
    from rapidnet_b200 import cabi
    fmodes = {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}
    rows, summary = [], []
    for case in args.configs.split(","):
        for iters in [int(x) for x in args.iters.split(",")]:
            prob, slot = load_problem(case, iters)
            c, fc = prob.config, prob.forecast
            ref, ref_s = run_reference(prob, slot)
            o64 = Oracle(prob, L=ref["L"], Lhat=ref["Lhat"], precision="f64")
            o64.factor_step(); o64.update_state(); o64.eliminate(fc.demand[slot], fc.prices[slot]); o64.apg(iters)
            f64 = {lab: o64.get(on).astype(np.float64) for lab, _, _, on in VARS}
            f64["u0"] = o64.get("U")[: prob.network.nu].astype(np.float64)
            o64.close()
            for fm in args.factors.split(","):
                s = cabi.Solver(prob)
                s.set_modes(cabi.SWEEP_PERSISTENT, fmodes[fm])
                s.factor_step()
                u0 = s.control_action(c.current_x, c.prev_u, c.prev_demand, fc.demand[slot], fc.prices[slot], iters)
                persistent = s.info().sweep_mode == cabi.SWEEP_PERSISTENT
                worst_ratio, worst_agree, acc_ok, agree_ok, limited = 0.0, 0.0, True, True, False
                pairs = []
                for lab, gname, rname, _ in VARS + [("u0", None, "u0", None)]:
                    ours = u0 if lab == "u0" else s.read(gname)
                    e_or, e_o64, e_r64 = rel_err(ours, ref[rname]), rel_err(ours, f64[lab]), rel_err(ref[rname], f64[lab])
                    a_ok = e_o64 <= ACC_FACTOR * e_r64 + ACC_ABS
                    lim = e_r64 > RTOL / (1.0 + ACC_FACTOR)
                    g_ok = e_or <= (RTOL if not lim else (1.0 + ACC_FACTOR) * e_r64)
                    rows.append({"config": case, "iterations": iters, "factors": fm, "variable": lab, "ours_vs_ref": e_or,
                                 "ours_vs_f64": e_o64, "ref_vs_f64": e_r64, "accuracy_gate": bool(a_ok),
                                 "agreement_gate": bool(g_ok), "floor_limited": bool(lim)})
                    if max(e_o64, e_r64) > ACC_ABS:   # cells at the rounding floor are not ranked
                        worst_ratio = max(worst_ratio, e_o64 / max(e_r64, 1e-12))
                    pairs.append((e_o64, e_r64))
                    worst_agree = max(worst_agree, e_or)
                    acc_ok &= a_ok; agree_ok &= g_ok; limited |= lim
                u0row = rows[-1]
                med = median_ratio(pairs)
                acc_ok &= med <= ACC_MEDIAN
                summary.append({"config": case, "iterations": iters, "factors": fm, "nodes": int(prob.tree.nodes),
                                "persistent_kernel": bool(persistent), "u0_ours_vs_ref": u0row["ours_vs_ref"],
                                "u0_ours_vs_f64": u0row["ours_vs_f64"], "u0_ref_vs_f64": u0row["ref_vs_f64"],
                                "worst_ours_vs_ref": worst_agree, "worst_accuracy_ratio": worst_ratio, "median_accuracy_ratio": med,
                                "accuracy_gate": bool(acc_ok), "agreement_gate": bool(agree_ok), "floor_limited": bool(limited),
                                "reference_seconds": round(ref_s, 2)})
                print(f"{case:6s} it={iters:4d} {fm:6s}: u0 ours-ref {u0row['ours_vs_ref']:.2e} ours-f64 {u0row['ours_vs_f64']:.2e} "
                      f"ref-f64 {u0row['ref_vs_f64']:.2e} | worst ours-ref {worst_agree:.2e}, worst accuracy ratio {worst_ratio:.2f} median {med:.2f} "
                      f"| accuracy {'ok' if acc_ok else 'FAIL'} agreement {'ok' if agree_ok else 'FAIL'}"
                      f"{' (floor-limited)' if limited else ''}", flush=True)
                s.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"gates": {"accuracy": f"ours_vs_f64 <= {ACC_FACTOR} * ref_vs_f64 + {ACC_ABS} per variable, and the median over the variables of ours_vs_f64 / ref_vs_f64 <= {ACC_MEDIAN}",
                         "agreement": f"ours_vs_ref <= {RTOL}, or <= {1 + ACC_FACTOR} * ref_vs_f64 when ref_vs_f64 > {RTOL}/{1 + ACC_FACTOR} (floor-limited)"},
               "summary": summary, "rows": rows}, open(args.out + ".json", "w"), indent=1)
    with open(args.out + ".md", "w") as f:
        f.write("# Parity: ours vs the reference build vs the double-precision trajectory\n\n"
                "Generated by `tools/parity_table.py` on a B200 (same box for all three).  Errors are norm-wise relative.  "
                "`ref` = the reference's own CUDA/cuBLAS build (`oracle/_ref/ref_driver`), `f64` = the oracle port compiled in double, "
                "fed the reference's null-space basis.  Gates: `tests/refcompare.py`.\n\n## Summary (u0 and the worst variable)\n\n"
                "| config | nodes | iterations | factors | u0 ours-ref | u0 ours-f64 | u0 ref-f64 | worst ours-ref | worst (ours-f64)/(ref-f64) | median (ours-f64)/(ref-f64) | accuracy gate | agreement gate |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in summary:
            f.write(f"| {r['config']} | {r['nodes']} | {r['iterations']} | {r['factors']} | {r['u0_ours_vs_ref']:.2e} | {r['u0_ours_vs_f64']:.2e} | "
                    f"{r['u0_ref_vs_f64']:.2e} | {r['worst_ours_vs_ref']:.2e} | {r['worst_accuracy_ratio']:.2f} | {r['median_accuracy_ratio']:.2f} | "
                    f"{'ok' if r['accuracy_gate'] else 'FAIL'} | {'ok' if r['agreement_gate'] else 'FAIL'}{' (floor-limited)' if r['floor_limited'] else ''} |\n")
        f.write("\n## Per variable\n\n| config | iterations | factors | variable | ours-ref | ours-f64 | ref-f64 | accuracy | agreement |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['config']} | {r['iterations']} | {r['factors']} | {r['variable']} | {r['ours_vs_ref']:.2e} | {r['ours_vs_f64']:.2e} | "
                    f"{r['ref_vs_f64']:.2e} | {'ok' if r['accuracy_gate'] else 'FAIL'} | "
                    f"{'ok' if r['agreement_gate'] else 'FAIL'}{' (floor-limited)' if r['floor_limited'] else ''} |\n")
    print("wrote", args.out + ".md")


if __name__ == "__main__":
    main()
