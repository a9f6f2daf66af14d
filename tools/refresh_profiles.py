"""Turns the artefacts of one tools/gpu_round.sh call (gpurun_out/<tag>_*) into the tracked summaries under profiles/<name>_*:
bench lines, the ncu launch list and the full-capture summary of the int8 Gram kernels (tools/ncu_summary.py + tensor-pipe counters).
usage: python tools/refresh_profiles.py <tag> <name>      e.g.  r01r r01_o"""
import csv
import io
import re
import shutil
import subprocess
import sys

tag, name = sys.argv[1], sys.argv[2]
shutil.copy(f"gpurun_out/{tag}_bench.json", f"profiles/{name}_bench_cfg2.json")
shutil.copy(f"gpurun_out/{tag}_ref.json", f"profiles/{name}_bench_cfg2_reference_arm.json")

rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches.csv")) if len(r) > 5 and r[0].isdigit()]
with open(f"profiles/{name}_launches_cfg2.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 : python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e\n"
            "# (cold-cache, serialised launches: compare SHARES, not absolutes).  id, kernel, duration_ns.  One update = prologue, affine (aux branch),\n"
            "# slice, gram_i8<5,96,1,64>, gram_i8<6,80,1,128> (no-op: the certificate picked 5 planes), gram_tma (fp64 companion, no-op)\n")
    for r in rows:
        f.write(f"{r[0]},{r[4]},{int(float(r[-1]))}\n")

rep = f"gpurun_out/{tag}_prof_i8.ncu-rep"
base = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout.split("\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
ki = hdr.index("Kernel Name")
pat = re.compile(r"utcimma_src_int8_sparsity_off\.(avg\.pct_of_peak_sustained_elapsed|avg\.peak_sustained$|avg\.per_cycle_elapsed|max\.pct_of_peak_sustained_elapsed|"
                 r"min\.pct_of_peak_sustained_elapsed|sum$)|^sm__mem_tensor_cycles_active\.avg\.pct|^sm__pipe_tensor_cycles_active\.avg\.pct|^sm__cycles_elapsed\.max|"
                 r"^lts__throughput\.avg\.pct|^l1tex__throughput\.avg\.pct|^smsp__issue_active\.avg\.pct")
want = [i for i, h in enumerate(hdr) if pat.search(h)]
extra = {"kernel: " + r[ki]: [f"  {hdr[i]:<100s}{r[i]:>18s} {units[i]}" for i in want if not ("slice" in r[ki] and "tensor" in hdr[i])] for r in rr[2:]}
out, cur, seen = [], None, set()


def flush():
    for line in extra.get(cur, []):
        if line.split()[0] not in seen:
            out.append(line)


for line in base:
    if line.startswith("#"):
        continue
    if line.startswith("kernel:"):
        flush(); cur = line; seen = set(); out.append(line); continue
    if line.startswith("  "):
        if cur and "slice" in cur and "tensor" in line:
            continue
        seen.add(line.split()[0])
    if line.strip():
        out.append(line)
flush()
with open(f"profiles/{name}_gram_i8_ncu_full.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:gram_i8 -s 6 -c 3 : python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e\n"
            "# cfg2 (n = k = 4096): gram_i8_slice_kernel<256,ring,full>, gram_i8_kernel<5,96,1,64> (64-byte stages, accumulators released before the stores),\n"
            "# <6,80,1,128> (128-byte split ring; no-op here).  Tensor-pipe evidence for tcgen05 kind::i8 = sm__ops_path_tensor_op_utcimma_src_int8*\n")
    f.write("\n".join(out) + "\n")
print("profiles refreshed from", tag)
