#!/usr/bin/env python
"""Turns the raw outputs of one GPU run (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py [tag]       # tag defaults to r01

Inputs (written by the command in profiles/README.md):
    gpurun_out/<tag>_bench_c2.json   bench.py JSON line
    gpurun_out/<tag>_launches.csv    ncu --metrics gpu__time_duration.sum launch list
    gpurun_out/<tag>_full.ncu-rep    ncu --set full capture of rgb_strip / lay_tile / pass2
"""
import csv, json, os, re, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout

# 1. bench line
b = os.path.join(G, f"{tag}_bench_c2.json")
if os.path.exists(b):
    open(os.path.join(P, f"{tag}_bench_c2.json"), "w").write(open(b).read())

# 2. launch list: keep the CSV and add the share of each kernel in one step
l = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(l):
    txt = open(l).read()
    open(os.path.join(P, f"{tag}_launches.csv"), "w").write(txt)
    rows = [r for r in csv.reader(txt.split("\n")) if len(r) > 14 and r[0].isdigit()]
    tot = defaultdict(float); cnt = Counter()
    for r in rows:
        k = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        tot[k] += float(r[14]); cnt[k] += 1
    s = sum(tot.values())
    with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
        f.write("share of the profiled launches (cold-cache, serialised; the SHARE is what must agree with bench.py)\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"{v / s * 100:6.2f}%  {v / cnt[k] / 1e3:8.1f} us/launch  x{cnt[k]:3d}  {k}\n")

# 3. full capture: key metrics, stalls, traffic
rep = os.path.join(G, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(raw.split("\n")))
    hdr = rows[0]
    want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__thread_inst_executed_per_inst_executed.ratio']
    stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    units = rows[1]
    traffic = {"workload": "c2: 16x256x512 K=20 f32 fwd+bwd", "source": f"ncu --set full --clock-control none ({tag}_full.ncu-rep)", "kernels": {}}
    with open(os.path.join(P, f"{tag}_ncu_full_kernels.txt"), "w") as f:
        for r in rows[2:]:
            if len(r) < len(hdr): continue
            name = r[hdr.index('Kernel Name')]
            f.write(f"==== {name}\n")
            for w in want:
                if w in hdr: f.write(f"  {w:78s} {r[hdr.index(w)]} {units[hdr.index(w)]}\n")
            st = sorted(((float(r[hdr.index(h)] or 0), h.split('issue_stalled_')[1].split('_per_')[0]) for h in stall), reverse=True)
            f.write("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st if v >= 0.05) + "\n")
            key = re.sub(r"[<(].*", "", name).replace("void ", "").strip()
            def val(m, scale=1.0):
                u = units[hdr.index(m)]
                v = float(r[hdr.index(m)])
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1) * scale
            traffic["kernels"][key] = {
                "dram_bytes_read": val('dram__bytes_read.sum'), "dram_bytes_write": val('dram__bytes_write.sum'),
                "gpu_time_us": float(r[hdr.index('gpu__time_duration.sum')]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[hdr.index('gpu__time_duration.sum')], 1),
                "lsu_data_pipe_pct": float(r[hdr.index('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')]),
                "issue_active_pct": float(r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]),
                "dram_pct_of_peak": float(r[hdr.index('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')]),
                "warp_instructions": float(r[hdr.index('smsp__inst_executed.sum')]),
                "shared_wavefronts": float(r[hdr.index('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')]),
            }
    json.dump(traffic, open(os.path.join(P, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    # shared-memory wavefronts per opcode + hottest stall sites
    cub, sas = os.path.join(G, "k_prof.cubin"), os.path.join(G, "k_prof.sass")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin", "-o", cub,
                    os.path.join(ROOT, "video-layout-generation_b200", "csrc", "vlg_api.cu")], check=True)
    open(sas, "w").write(run(["nvdisasm", "-g", "-c", cub]))
    with open(os.path.join(P, f"{tag}_ncu_source_hotspots.txt"), "w") as f:
        for kre, ksub, mainf in (("rgb_strip", "rgb_strip_kernelIfLb1ELb1", "vlg_rgb.cuh"), ("lay_tile", "lay_tile_kernelIfLi20ELb1ELb0", "vlg_laytile.cuh"),
                                 ("pass2_rec_kernel", "pass2_rec_kernelIfLi20E", "vlg_pass2.cuh")):
            src = os.path.join(G, f"{tag}_{kre}_src.csv")
            open(src, "w").write(run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"]))
            f.write(f"######## {kre}\n")
            f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_smem.py"), src, "65536"]))
            f.write("-- instructions with the most stall samples (share, kernel line, inlined location, SASS, top stall reasons)\n")
            f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), src, sas, ksub, mainf, "12"]))

# 4. SASS mnemonics that prove the Blackwell paths (TMA, mbarrier, packed fp32x2, REDUX)
lib = os.path.join(ROOT, "video-layout-generation_b200", "libvlg_b200.so")
if os.path.exists(lib):
    sass = run(["cuobjdump", "-sass", lib])
    c = Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass, re.M):
        op = m.group(1)
        for k in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "CREDUX", "REDUX", "LDGSTS", "UTMACMDFLUSH", "FENCE"):
            if op.startswith(k): c[k] += 1
    with open(os.path.join(P, f"{tag}_sass_blackwell_ops.txt"), "w") as f:
        f.write("static SASS counts in libvlg_b200.so (cuobjdump -sass): TMA tensor loads (UTMALDG), bulk copies (UBLKCP),\n"
                "mbarrier (SYNCS), packed fp32x2 (FFMA2/FADD2/FMUL2), warp reductions (CREDUX/REDUX), cp.async (LDGSTS)\n")
        for k, v in c.most_common(): f.write(f"{v:7d} {k}\n")
print("profiles written for", tag)
