"""Static SASS mnemonic counts of the main kernels (runs here, no GPU):  python tools/sass_mnemonics.py > profiles/..."""
import collections
import re
import subprocess
import sys

LIB = "steered-mixture-of-experts_b200/libsmoe_b200.so"
WANT = ("loss_kernelILi3E", "forward_kernelILi2ELi3ELb0ELb0", "grad_finalize_kernelILi2ELi3ELb1", "backward_kernelILi2ELi3ELb0ELb0",
        "backward_kernelILi2ELi3ELb0ELb1", "bwd_plan_kernelILi2ELi3", "halo_pull_kernel", "xchg_publish_kernel",
        "ssim2d_tile_kernelILi3ELb0", "ssim2d_tile_kernelILi3ELb1", "sl_adj_apply_tileILi2ELi3")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("Static SASS mnemonic counts (cuobjdump -sass libsmoe_b200.so, sm_100a) of the round-2 kernels (tools/sass_mnemonics.py).")
print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS.* = mbarrier ops, LDGSTS = cp.async (4-byte asynchronous copies), MUFU.EX2 = ex2.approx,")
print("LD/ST .SYS / .STRONG.SYS = peer-memory exchange; the ATOM* entries are integer tickets, counters and bitmasks (no float atomics).\n")
cur, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if any(w in m.group(1) for w in WANT) else None
        if cur:
            counts[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            keep = op if op.startswith(("LDG", "STG", "LD.", "ST.", "LDS", "STS", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "ATOM", "RED", "REDUX")) else op.split(".")[0]
            counts[cur][keep] += 1
for fn, c in counts.items():
    print(fn)
    print("  " + ", ".join(f"{k}:{v}" for k, v in c.most_common(44)))
    special = {k: v for k, v in c.items() if k.startswith(("UBLKCP", "SYNCS", "LDGSTS", "ATOM", "RED"))}
    print("  of note: " + (", ".join(f"{k}:{v}" for k, v in sorted(special.items())) or "-") + "\n")
