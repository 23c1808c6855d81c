"""Time the device-resident step for every library variant in _variants (run on the GPU box).
usage: python tools/tune_run.py [cells] [records] [chunk ...]"""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cells = sys.argv[1] if len(sys.argv) > 1 else "100000"
recs = sys.argv[2] if len(sys.argv) > 2 else "365"
chunks = sys.argv[3:] if len(sys.argv) > 3 else ["0"]
for lib in sorted(glob.glob(os.path.join(ROOT, "_variants", "lib_*.so"))):
  for ch in chunks:
    env = dict(os.environ, VICRES_LIB=lib, VICRES_TIME_KERNELS=os.environ.get("VICRES_TIME_KERNELS", "1"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "2", "--no-cpu-baseline",
                        "--e2e-steps", "1", "--no-forcing-prep", "--cells", cells, "--block", recs, "--chunk", ch], env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print("%-24s chunk %3s value %.4e  e2e %.4e  ms/step %.1f  k_units %.3f ms  sm %s MHz" % (os.path.basename(lib), ch, j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["clocks"]["sm_mhz"]),
              "| snow share %.3f us/record %s" % (j["step_kernels"]["snow_unit_record_share"],
                                                 {k: round(v, 1) for k, v in (j["step_kernels"]["us_per_record"] or {}).items()}), flush=True)
    except Exception as e:
        print(os.path.basename(lib), "FAILED", r.stderr[-400:], flush=True)
