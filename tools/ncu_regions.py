#!/usr/bin/env python3
"""Summarise an ncu report of one kernel by source region (development aid).

  python tools/ncu_regions.py gpurun_out/prof.ncu-rep k_fit_startsILb1 [libabfit.so that was profiled]

Joins `ncu --page source --csv` (per-SASS-instruction samples / executed counts) with
`nvdisasm --print-line-info` of the kernel cubin inside the profiled libabfit.so, and prints
instruction share, stall-sample share and average active threads per region of the device code,
plus the headline raw metrics.  Needs no GPU."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "alphabeta-rs_b200/csrc/abfit_kernels.cu")
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "alphabeta-rs_b200/libabfit.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "abfit_kernels.sm_100a.cubin", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = os.path.join(tmp, "abfit_kernels.sm_100a.cubin")
sass = subprocess.check_output(["nvdisasm", "--print-line-info", cubin], text=True).split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
line_of, cur = {}, None
for l in sass[start + 1:]:
    if l.startswith(".text.") or "//----" in l and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))

raw = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True).split("\n")))
hdr, vals = raw[0], raw[2]
for want in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
             "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
             "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if want in hdr:
        print(f"{want:70s} {vals[hdr.index(want)]}")
for h, v in zip(hdr, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v) > 0.05:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {float(v):.3f}")

rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True).split("\n")))
h2 = rows[1]
data = [r for r in rows[2:] if len(r) == len(h2)]
ia, isamp, iex, ith = h2.index("Address"), h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
base = int(data[0][ia], 16)


def region(k):
    if k is None:
        return "none"
    f, l = k
    if f == "abfit_model.cuh":
        txt = open(os.path.join(os.path.dirname(src), f)).read().split("\n")
        # region = enclosing function name
        for j in range(l - 1, -1, -1):
            m = re.match(r"^(?:template.*\n)?__device__ __forceinline__ \S+ (\w+)\(", txt[j])
            if m:
                return "model:" + m.group(1)
        return "model:?"
    if f == "abfit_nm.cuh":
        return "nm"
    return f


tot = [0, 0, 0]
reg = collections.defaultdict(lambda: [0, 0, 0])
mismatch = 0
for r in data:
    off = int(r[ia], 16) - base
    cur, ins = line_of.get(off, (None, ""))
    if ins.split()[:1] != r[h2.index("Source")].split()[:1]:
        mismatch += 1
    a = reg[region(cur)]
    for i, c in enumerate((isamp, iex, ith)):
        a[i] += int(r[c]); tot[i] += int(r[c])
print(f"SASS instructions {len(data)} (opcode mismatches vs local build: {mismatch}); executed {tot[1]:.3e}; avg threads {tot[2] / tot[1]:.1f}")
for k, a in sorted(reg.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} inst {a[1] / tot[1] * 100:5.1f}%  samples {a[0] / tot[0] * 100:5.1f}%  avg_thr {a[2] / max(a[1], 1):5.1f}")
