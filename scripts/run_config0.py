"""Record configs[0] (tests/config0.py) for both loops: profiles/r2_config0_trace.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import config0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
f = config0.objective_function()
out = {"config": "test_1a.py shape: d=4, m=5, 6^4-grid GP-sample attributes, EI-CF, fixed_hyps=True, seed 0, %d iterations" % iters}
for side in ("cuda", "cpu"):
    t0 = time.time()
    out[side] = config0.run(side, f, iters, seed=0)
    out[side]["wall_s"] = time.time() - t0
Xg, Xc = np.array(out["cuda"]["suggested_points"]), np.array(out["cpu"]["suggested_points"])
out["max_abs_point_difference_per_iteration"] = np.abs(Xg - Xc).max(axis=1).tolist()
out["best_value_trace_abs_difference"] = np.abs(np.array(out["cuda"]["best_value_trace"]) - np.array(out["cpu"]["best_value_trace"])).tolist()
path = os.path.join(ROOT, "gpurun_out", "r2_config0_trace.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps({k: out[k] for k in ("max_abs_point_difference_per_iteration", "best_value_trace_abs_difference")}))
print("wall: cuda %.1f s, cpu %.1f s" % (out["cuda"]["wall_s"], out["cpu"]["wall_s"]))
