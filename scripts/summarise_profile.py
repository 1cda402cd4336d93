"""Copy one round_end.sh run (gpurun_out/<name>_*) into profiles/<tag>_* and print the summary that profiles/README.md quotes:
stall shares and SASS mnemonic counts of the `ncu --set full` capture, DRAM traffic, per-kernel share of the launch list."""
import collections, csv, re, shutil, subprocess, sys
name, tag = sys.argv[1], sys.argv[2]
src = "gpurun_out/%s_source.csv" % name
with open(src, "w") as fo:
    subprocess.run(["ncu", "-i", "gpurun_out/%s_persistent.ncu-rep" % name, "--page", "source", "--csv"], stdout=fo, stderr=subprocess.DEVNULL)
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {c: i for i, c in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) >= len(h) and r[0].startswith("0x")]
def f(r, c):
    try: return int(float(r[ix[c]].replace(",", "")))
    except Exception: return 0
cols = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_lg", "stall_mio", "stall_math", "stall_membar"]
body.sort(key=lambda r: -f(r, "# Samples"))
with open("profiles/%s_persistent_ncu_source_top120.csv" % tag, "w", newline="") as fo:
    w = csv.writer(fo); w.writerow(["Address", "Source", "# Samples", "Instructions Executed"] + cols)
    for r in body[:120]: w.writerow([r[0], r[1], f(r, "# Samples"), f(r, "Instructions Executed")] + [f(r, c) for c in cols])
tot = sum(f(r, "# Samples") for r in body)
print("stalls:", {c: "%.1f%%" % (100 * sum(f(r, c) for r in body) / tot) for c in cols})
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", r[1].strip()).split()[0].split(".")[0] for r in body)
print("sass:", {k: ops[k] for k in ["UBLKCP", "LDGSTS", "SYNCS", "CCTL", "DFMA", "LDS", "LDG", "STG", "BAR", "SHFL"]})
raw = list(csv.reader(open("gpurun_out/%s_raw.csv" % name)))
d = dict(zip(raw[0], zip(raw[1], raw[2])))
for k in ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread"]:
    print(k, d.get(k))
for a, b in [("raw.csv", "persistent_ncu_raw.csv"), ("launches.csv", "launches_persistent_bench.csv"), ("bench.log", "bench.json"), ("bench_ref.log", "bench_reference_arm.json")]:
    shutil.copy("gpurun_out/%s_%s" % (name, a), "profiles/%s_%s" % (tag, b))
L = list(csv.reader(l for l in open("gpurun_out/%s_launches.csv" % name) if l.startswith('"')))
k, v = L[0].index("Kernel Name"), L[0].index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in L[1:]:
    n = re.sub(r"\(.*", "", r[k]).replace("void ", ""); agg[n][0] += 1; agg[n][1] += float(r[v].replace(",", ""))
T = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:6]: print("%-32s %4d launches %10.3f ms %6.2f%%" % (n, a[0], a[1] / 1e6, 100 * a[1] / T))
