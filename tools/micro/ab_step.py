"""A/B the cfg2 train step between libclipgp.so and libclipgp_ts.so on the same box (kernel experiments)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for rnd in range(4):
    for which, lib in (("main", ""), ("alt ", "libclipgp_ts.so")):
        env = dict(os.environ); env["CLIPGP_LIB"] = lib
        if os.environ.get("AB_ENV") and lib:            # A/B an engine toggle instead of a second library: AB_ENV=CLIPGP_NO_FUSE_OPT
            env["CLIPGP_LIB"] = ""; env[os.environ["AB_ENV"]] = "1"
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "micro", "time_step.py"), "tf32", "300"], capture_output=True, text=True, env=env)
        print(which, (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1])
