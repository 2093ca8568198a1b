#!/usr/bin/env python
"""scripts/sass_evidence.py -- counts of selected SASS mnemonics per kernel of the built library (cuobjdump -sass): DMMA (fp64
tensor-core MMA), UBLKCP (cp.async.bulk = TMA bulk copy), SYNCS.* (mbarrier), FENCE.VIEW.ASYNC (proxy fence), 256-bit stores ...
Writes profiles/<tag>_sass_evidence.txt.  CPU only."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bluerov2_b200", "lib", "libacados_ocp_solver_bluerov2.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "x"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur:
        cnt[cur][m.group(1)] += 1
want = ["DMMA.8x8x4", "LDGSTS.E.BYPASS.128", "LDGDEPBAR", "UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "SYNCS.PHASECHK.TRANS64", "FENCE.VIEW.ASYNC.S",
        "CCTL.E.PF2", "STG.E.ENL2.256", "LDS.128", "DFMA", "SHFL.IDX", "SHFL.BFLY", "MUFU.RCP64H", "MUFU.RSQ64H", "CREDUX"]
names = {"pdas_kernel": "pdas_kernel", "ipm_kernel": "ipm_kernel", "eskf_predict": "eskf_predict_kernel", "eskf_update": "eskf_update_kernel",
         "exchange_kernel": "exchange_kernel", "linearize_kernel": "linearize_kernel", "ekf_kernel": "ekf_kernel", "rls_kernel": "rls_kernel", "plant_kernel": "plant_kernel",
         "yaw_unwrap_kernel": "yaw_unwrap_kernel"}
lines = ["# SASS evidence: cuobjdump -sass bluerov2_b200/lib/libacados_ocp_solver_bluerov2.so (sm_100a), selected mnemonics per kernel",
         "# DMMA.8x8x4 = mma.sync.m8n8k4.f64; UBLKCP.S.G = cp.async.bulk global->shared (TMA bulk copy); SYNCS.* = mbarrier arrive.expect_tx /",
         "# try_wait / test_wait; FENCE.VIEW.ASYNC = fence.proxy.async; CCTL.E.PF2 = prefetch.global.L2; STG.E.ENL2.256 = st.global.v4.f64;",
         "# CREDUX = __reduce_{max,min}_sync; LDGSTS.E.BYPASS.128 = cp.async.cg 16 B (global -> shared, L1 bypassed); LDGDEPBAR = cp.async.commit_group"]
for k in sorted(cnt, key=lambda n: [v for key, v in names.items() if key in n] or ["~"]):
    nm = [v for key, v in names.items() if key in k]
    if not nm:
        continue
    c = cnt[k]
    lines.append(f"== {nm[0]}   ({sum(c.values())} instructions)")
    for w in want:
        tot = sum(v for op, v in c.items() if op == w or op.startswith(w + "."))
        if tot:
            lines.append(f"   {tot:6d} {w}")
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
lines += ["", "arch:"] + [l for l in arch.splitlines() if "cubin" in l]
path = os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.txt")
open(path, "w").write("\n".join(lines) + "\n")
print(path)
