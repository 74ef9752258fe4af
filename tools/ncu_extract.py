"""Turns an .ncu-rep of the headline launch into the committed evidence under profiles/:
   python tools/ncu_extract.py <report.ncu-rep> <out_prefix>
writes <out_prefix>_raw.csv (ncu --page raw), prints the metrics DESIGN.md / the ncu summary quote, and a
per-region breakdown of the warp-state samples from the source page (producer wait, fused first k-iteration,
main loop, rest) so the "where the remaining percent goes" paragraph can be re-derived."""
import csv
import io
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(prefix + "_raw.csv", "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} [{units[i]}] = {vals[i]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = lambda r: int(r[ix["# Samples"]])
total = sum(S(r) for r in data)
dm = [n for n, r in enumerate(data) if "DMMA" in r[ix["Source"]]]
clusters, st, pr = [], dm[0], dm[0]
for d in dm[1:]:
    if d - pr > 60:
        clusters.append((st, pr))
        st = d
    pr = d
clusters.append((st, pr))
main = clusters[-1]
top = main[0]
while "TRYWAIT" not in data[top][ix["Source"]]:
    top -= 1
top -= 4
fused = None
if len(clusters) > 1:
    ftop = clusters[0][0]
    while "TRYWAIT" not in data[ftop][ix["Source"]]:
        ftop -= 1
    fused = (ftop - 4, clusters[-2][1])   # from its own barrier wait to its last DMMA
# the producer's wait: the single hottest branch outside the consumer clusters
prod = max((n for n in range(len(data)) if n < clusters[0][0] - 200), key=lambda n: S(data[n]))
rng = lambda a, b: sum(S(r) for r in data[a:b + 1])
print(f"warp-state samples: total {total}")
print(f"  producer thread waiting for a free slot (one branch): {S(data[prod])} ({100 * S(data[prod]) / total:.1f} %)")
if fused:
    print(f"  fused first k-iteration of a tile (rows {fused[0]}-{fused[1]}): {rng(*fused)}")
print(f"  main loop incl. its barrier wait (rows {top}-{main[1]}): {rng(top, main[1])}; per k-iteration of 128 DMMAs: "
      f"{rng(top, main[1]) / max(1, int(data[main[0]][ix['Instructions Executed']])) * 1:.6f} samples per execution")
tries = [int(r[ix["Instructions Executed"]]) for r in data[top:main[0]] if "TRYWAIT" in r[ix["Source"]]]
iters = int(data[main[0]][ix["Instructions Executed"]])
if tries:
    print(f"  mbarrier try_wait executions per main-loop iteration: {tries[0] / iters:.2f} ({tries[0]} / {iters})")
if fused:
    fe = max(int(data[n][ix["Instructions Executed"]]) for n in range(clusters[0][0], clusters[0][1]) if "DMMA" in data[n][ix["Source"]])
    print(f"  fused iteration / plain iteration, samples per execution: {rng(*fused) / fe / (rng(top, main[1]) / iters):.3f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h[6:]: sum(int(r[ix[h]]) for r in data[top:main[1] + 1]) for h in stalls}
allm = sum(tot.values())
print("  main-loop warp states: " + ", ".join(f"{k} {100 * v / allm:.1f} %" for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v * 200 > allm))
