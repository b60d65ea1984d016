#!/usr/bin/env python3
"""profiles/r02_peaks.json: the denominators bench.py's roofline uses, measured on the device with clocks sampled during the probes:
FP64 issue rates (DADD, DMUL, DFMA; 1e12 thread-instructions / s), the DFMA flop rate, device copy bandwidth.   usage: python tools/peaks.py out.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xfluids_b200 import capi

if __name__ == "__main__":
    s = bench.ClockSampler(0)
    s.start()
    t0 = time.time()
    issue = [capi.measure_fp64_issue(0) for _ in range(3)]
    dfma, copy = capi.measure_peaks(0)
    clocks = s.stop()
    best = [max(x[i] for x in issue) for i in range(3)]
    out = {"device": 0, "fp64_issue_tinst_per_s": {"DADD": best[0], "DMUL": best[1], "DFMA": best[2]}, "dfma_tflops": dfma, "copy_gbs": copy,
           "nominal_fp64_tflops": 37.0, "seconds": time.time() - t0, "clocks": clocks,
           "how": "xf_measure_fp64_issue / xf_measure_peaks (csrc/xf_capi.cu): 8 independent chains per thread, 8 x 256-thread blocks per SM, best of 5; copy: 2 GiB double2 copy, read + write bytes"}
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(out))
