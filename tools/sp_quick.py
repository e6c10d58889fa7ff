"""Development aid: one-line timing of the single-phase pipeline at the default settings (HLL and LxF), 2^23 cells."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyperelasticsolver_b200 as H
from tools.sp_pipeline_bench import run
if __name__ == "__main__":
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    for flux in ("hll", "lxf"):
        run(logn, 20, flux, {})
