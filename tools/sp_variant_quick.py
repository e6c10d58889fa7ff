"""Development aid: one line per flux for the library / environment this process was started with (HYPERELASTIC_B200_LIB,
HS_SP_L2PROMO, ...): single-phase step, 2^logn cells, best of 3 x 20 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sp_pipeline_bench import run
if __name__ == "__main__":
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    tag = {"lib": os.path.basename(os.environ.get("HYPERELASTIC_B200_LIB", "default")), "promo": os.environ.get("HS_SP_L2PROMO", "128")}
    print(tag, flush=True)
    run(logn, 20, "hll", {})
    if len(sys.argv) > 2:
        run(logn, 20, "lxf", {})
