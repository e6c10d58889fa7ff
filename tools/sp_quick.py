"""Development aid: one-line timings of the single-phase pipeline (HLL and LxF), 2^23 cells; extra arguments KEY=VALUE
are environment settings to compare against the default (e.g. HS_SP_SINGLE=0 HS_SP_TILES=4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sp_pipeline_bench import run
if __name__ == "__main__":
    envs = [{}] + [dict([a.split("=", 1)]) for a in sys.argv[1:]]
    for _ in range(2):
        for env in envs:
            run(23, 20, "hll", env)
    run(23, 20, "lxf", {})
