"""prw_kernel sweeps: library build (RBG_B200_LIB, a file name in lib/) x pool size (RBG_PRW_M) x threads per CTA (RBG_PRW_THREADS); one
subprocess per setting.  (RBG_PRW_CHAIN belonged to the round-2 experiment recorded in DESIGN.md K1 and profiles/r02e_prw_chain_sweep*.log.)"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libdir = os.path.join(root, "routing-board-generation_b200", "lib")
settings = []
for lib in ("librbg_b200.so", "librbg_c5.so"):
    for chain in ("1", "0"):
        for (m, t) in ((0, 0), (64, 128), (128, 256), (96, 192), (32, 128)):
            if chain == "0" and (m, t) != (0, 0):
                continue
            settings.append({"RBG_B200_LIB": os.path.join(libdir, lib), "RBG_PRW_CHAIN": chain, "RBG_PRW_M": str(m), "RBG_PRW_THREADS": str(t)})
if len(sys.argv) > 1:
    settings = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
    for s in settings:
        if "RBG_B200_LIB" in s:
            s["RBG_B200_LIB"] = os.path.join(libdir, s["RBG_B200_LIB"])
for env in settings:
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_prw.py"), "20"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    tag = " ".join(f"{k.replace('RBG_', '')}={os.path.basename(v)}" for k, v in env.items())
    print(tag, "|", " | ".join(l.split(":")[0].replace("prw ", "") + ":" + l.split("ms")[1].split("boards/s")[0] for l in r.stdout.strip().splitlines()) or r.stderr[-400:], flush=True)
