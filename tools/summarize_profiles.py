"""Turn gpurun_out/<tag>_* artefacts into the tracked summaries under profiles/:
  <tag>_launches_by_kernel.txt  per-kernel share of one bench.py step (ncu gpu__time_duration launch list)
  <tag>_launches.csv            the raw launch list (kernel, grid, block, ns)
  <tag>_ncu_top.txt             key `ncu --set full` metrics of the top kernels (from <tag>_top.ncu-rep)
  <tag>_bench.json / _bench_ref.json / _profile_step.log / _pytest_gpu.log   copied verbatim
usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

lc = os.path.join(G, tag + "_launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.defaultdict(lambda: [0, 0.0])
    with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
        f.write("kernel,grid,block,ns\n")
        for r in rows:
            name = re.sub(r"\(.*", "", r[ki])
            ns = float(r[vi].replace(",", ""))
            agg[name][0] += 1
            agg[name][1] += ns / 1e6
            f.write('"%s","%s","%s",%d\n' % (name, r[gi], r[bi], ns))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, tag + "_launches_by_kernel.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off, one timed step of\n"
                "# `python bench.py --steps 1 --warmup 3` (cold-cache, serialised: read SHARES)\n")
        f.write("launches %d  total %.3f ms\n" % (len(rows), tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%9.3f ms %5.1f%%  x%-4d %s\n" % (v[1], 100 * v[1] / tot, v[0], k))

for rep in sorted(f for f in os.listdir(G) if f.startswith(tag) and f.endswith(".ncu-rep")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled"]
    idx = []
    for w in want:
        m = [i for i, h in enumerate(hdr) if h == w] or [i for i, h in enumerate(hdr) if h.startswith(w)]
        idx += m[:12] if w.startswith("smsp__average_warp") or w.startswith("smsp__warp_issue") else m[:1]
    if rep == tag + "_top.ncu-rep":
        # DRAM traffic of the dominant kernel (first NT GEMM launch of tools/ncu_targets.py = FFN up-projection
        # + bias + GELU at the bench shape) -> bench.py's roofline.traffic
        import json
        ki = hdr.index("Kernel Name")
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for d in data:
            if "gemm_kernel<0, 0, 1>" in d[ki]:
                tb = float(d[ri]) * scale[units[ri]] + float(d[wi]) * scale[units[wi]]
                json.dump({"kernel": "gemm_kernel<NT> M=23968 N=3072 K=768 (+bias+GELU, 2 bf16 outputs)",
                           "dram_bytes_per_launch": tb, "source": "profiles/%s_ncu_top.txt (ncu --set full)" % tag},
                          open(os.path.join(P, "roofline_traffic.json"), "w"))
                break
    with open(os.path.join(P, rep.replace(".ncu-rep", ".txt").replace(tag + "_", tag + "_ncu_")), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (extract of %s)\n" % rep)
        for d in data:
            f.write("----\n")
            for i in idx:
                f.write("  %-80s %s %s\n" % (hdr[i][:80], d[i][:110], units[i]))

for suffix in ("_bench.json", "_bench_ref.json", "_profile_step.log", "_pytest_gpu.log", "_smi.txt"):
    src = os.path.join(G, tag + suffix)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, tag + suffix))
print("profiles/ updated for", tag)
