"""Re-pack the reference's shipped result pickles into small NumPy fixtures.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The pickles (`visualization/results_benchmark_{1st,2nd}_draft/*.pkl`, written by the reference's
own `save_results_pickle`, benchmark_SE3_tracking.py:272-327) hold the complete problem definition
and the reference solver's final trajectories and per-iteration histories.  They are the only
known-answer vectors the reference has.  Only the DDP entries
(`ms_*`, `ss_*`) are kept; the IPOPT baselines (`*_euc`) are out of scope.  Nothing here is
reference *source*; these are data produced by the reference.
"""
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/visualization"
OUT = os.path.dirname(os.path.abspath(__file__))

FILES = {
    # fixture name: (pickle, kind)
    "se3_n955_r1e-5": ("results_benchmark_2nd_draft/results_se3_tracking_drone_benchmark.pkl", "se3"),
    "se3_n955_r1e-4": ("results_benchmark_2nd_draft/results_se3_tracking_benchmark.pkl", "se3"),
    "se3_n120": ("results_benchmark_2nd_draft/results_se3_tracking_generate_benchmark.pkl", "se3"),
    "drone_n150": ("results_benchmark_2nd_draft/results_drone_racing_tracking_benchmark.pkl", "drone"),
    "so3_n249": ("results_benchmark_2nd_draft/results_so3_tracking_benchmark.pkl", "so3"),
    "pendulum_n80": ("results_benchmark_2nd_draft/results_pendulum_swingup_benchmark.pkl", "pendulum"),
    "draft1_se3_n955": ("results_benchmark_1st_draft/results_se3_tracking_benchmark.pkl", "se3"),
    "draft1_drone_n500": ("results_benchmark_1st_draft/results_drone_racing_tracking_benchmark.pkl", "drone"),
    "draft1_so3_n249": ("results_benchmark_1st_draft/results_so3_tracking_benchmark.pkl", "so3"),
}
# Not re-packed: results_benchmark_1st_draft/results_pendulum_swingup_benchmark.pkl.  It was written by an older
# revision of the reference library (like draft1_se3_n955, which is kept for its problem definition only): today's
# library code — and therefore the oracle — gives 41 multiple-shooting iterations against the file's 19, J_hist 8e-4
# relative off from the first iteration on.  The 2nd-draft pendulum file (pendulum_n80) replays to 1e-15.


def main():
    for name, (rel, kind) in FILES.items():
        with open(os.path.join(REF, rel), "rb") as f:
            d = pickle.load(f)
        p = d["prob"]
        out = {"kind": np.array(kind), "source": np.array(rel)}
        for key in ("J", "dt", "q_ref", "xi_ref", "Q", "P", "R"):
            out["prob_" + key] = np.asarray(p[key], dtype=float)
        for key in ("m", "length"):
            if key in p:
                out["prob_" + key] = np.asarray(p[key], dtype=float)
        out["prob_x0_q"] = np.asarray(p["x0"][0], dtype=float)
        out["prob_x0_xi"] = np.asarray(p["x0"][1], dtype=float)
        for entry in d:
            if not (entry.startswith("ms_") or entry.startswith("ss_")):
                continue
            e = d[entry]
            tag = entry[:2]
            out[tag + "_xs_q"] = np.stack([np.asarray(x[0], dtype=float) for x in e["xs"]])
            out[tag + "_xs_xi"] = np.stack([np.asarray(x[1], dtype=float).reshape(-1) for x in e["xs"]])
            out[tag + "_us"] = np.asarray(e["us"], dtype=float)
            for h in ("J_hist", "grad_hist", "defect_hist"):
                if h in e:
                    out[tag + "_" + h] = np.asarray(e[h], dtype=float).reshape(-1)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:22s} {os.path.getsize(path) / 1024:8.1f} KiB  keys={len(out)}")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; the committed .npz fixtures are the output of this script")
    main()
