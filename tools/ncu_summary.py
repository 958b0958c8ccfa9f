"""Summarise gpurun_out/<rep>.ncu-rep + launches csv into profiles/ (tracked). Usage: python tools/ncu_summary.py r1c [N G]"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 11
G = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(root, "gpurun_out", "prof_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
idx = [hdr.index(k) for k in keep if k in hdr]
out = os.path.join(root, "profiles", "%s_step_kernel_ncu_full.csv" % tag)
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
print(open(out).read())
def col(k):
    return [float(r[hdr.index(k)]) for r in rows[2:]]
def to_bytes(k):
    u = units[hdr.index(k)].lower()
    m = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    return [v * m for v in col(k)]
t = (sum(to_bytes('dram__bytes_read.sum')) + sum(to_bytes('dram__bytes_write.sum'))) / (len(rows) - 2)
tj_path = os.path.join(root, "profiles", "traffic.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
tj["N%d_G%d" % (N, G)] = {"dram_bytes_per_launch": t, "source": "profiles/%s_step_kernel_ncu_full.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches)" % (tag, len(rows) - 2)}
json.dump(tj, open(tj_path, "w"), indent=1)
print("traffic per launch: %.1f MB" % (t / 1e6))
lp = os.path.join(root, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lp):
    dst = os.path.join(root, "profiles", "%s_launches.csv" % tag)
    open(dst, "w").write(open(lp).read())
    rr = [r for r in csv.reader(open(lp)) if len(r) > 5]
    h = rr[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rr[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        if r[ui] == 'ns': v /= 1e3
        agg[r[ki][:64]][0] += 1; agg[r[ki][:64]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(root, "profiles", "%s_launches_summary.txt" % tag), "w") as f:
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            line = "%-66s n=%4d total=%9.1f us avg=%7.1f us share=%.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot)
            print(line); f.write(line + "\n")
