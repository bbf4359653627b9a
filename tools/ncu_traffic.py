"""Summarise an `ncu --set full` capture (raw page exported here with
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/<name>_raw.csv )
into profiles/<name>_summary.txt and profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
    python tools/ncu_traffic.py profiles/<name>_raw.csv <rows per launch> [K=2080]"""
import csv, json, os, re, sys

src = sys.argv[1]
rows = int(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2080
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]

with open(src, newline="") as f:
    rd = list(csv.reader(f))
hdr = next(i for i, r in enumerate(rd) if "Kernel Name" in r)
names, units = rd[hdr], rd[hdr + 1]
col = {n: i for i, n in enumerate(names)}
out, lines = {}, []
for r in rd[hdr + 2:]:
    if len(r) < len(names):
        continue
    kname = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", kname).replace("scfgp::emu::", "").replace("scfgp::", "")
    short = re.sub(r"^void\s+", "", short).replace("emu::", "")
    vals = {}
    for w in WANT:
        if w in col:
            try:
                vals[w] = float(r[col[w]].replace(",", ""))
            except ValueError:
                vals[w] = r[col[w]]
            vals[w + "#unit"] = units[col[w]]
    lines.append((short, r[col["ID"]], vals))

def gb(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return v * mult

txt = [f"ncu --set full --clock-control none, cfg4 geometry (K={K}), {rows} rows per launch; source {os.path.basename(src)}", ""]
seen = {}
for short, lid, v in lines:
    dr = gb(v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_read.sum#unit"))
    dw = gb(v.get("dram__bytes_write.sum", 0), v.get("dram__bytes_write.sum#unit"))
    dur = v.get("gpu__time_duration.sum", 0)
    du = v.get("gpu__time_duration.sum#unit")
    dur_us = dur * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(du, 1)
    rec = {"launch_id": lid, "duration_us": dur_us, "dram_read_bytes": dr, "dram_write_bytes": dw,
           "dram_bytes_per_launch": dr + dw,
           "tensor_pipe_active_pct": v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
           "l2_hit_pct": v.get("lts__t_sector_hit_rate.pct"),
           "warps_active_pct": v.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "registers": v.get("launch__registers_per_thread"),
           "dram_throughput_pct": v.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "fp64_pipe_active_pct": v.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")}
    seen.setdefault(short, []).append(rec)
    txt.append(f"{short:60s} id {lid:>4s}  {dur_us:9.1f} us  dram {(dr + dw) / 1e6:8.1f} MB  tensor {rec['tensor_pipe_active_pct']}%  "
               f"L2 hit {rec['l2_hit_pct']}%  warps {rec['warps_active_pct']}%  regs {rec['registers']}  dram% {rec['dram_throughput_pct']}  fp64% {rec['fp64_pipe_active_pct']}")
base = os.path.splitext(src)[0].replace("_raw", "")
open(base + "_summary.txt", "w").write("\n".join(txt) + "\n")
print("\n".join(txt))
# traffic table for bench.py: last captured launch of each kernel
tj_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
for short, recs in seen.items():
    m = re.match(r"(k_\w+)<(.*)>", short)
    key = short
    if m:
        args = [a.strip() for a in m.group(2).split(",")]
        base_name = m.group(1).replace("k_backward_i8_2p", "k_backward_i8")   # the two-phase variant of the same product
        key = f"{base_name}<{','.join(args[:2])}>" if base_name in ("k_backward_i8", "k_tgemm_i8") else f"{base_name}<{','.join(args)}>"
    r = recs[-1]
    tj[key] = dict(r, rows=rows, source=os.path.relpath(base + "_summary.txt", ROOT))
if "k_backward_i8<5,96>" in tj:
    d = 5
    a_bytes, c_bytes, phi_bytes, z_bytes = rows * 2 * K * d, K * 2 * K * d, rows * K * d, rows * (K // 2) * 8
    tj["k_backward_i8<5,96>"]["algorithmic_bytes"] = a_bytes + c_bytes + phi_bytes + z_bytes
    tj["k_backward_i8<5,96>"]["algorithmic_note"] = (
        f"A digits rows*2K*{d} ({a_bytes / 1e6:.0f} MB) + [2Abar|2sB] digits K*2K*{d} ({c_bytes / 1e6:.0f} MB) + Phi digits "
        f"re-read by the epilogue ({phi_bytes / 1e6:.0f} MB) + Zbar write ({z_bytes / 1e6:.0f} MB)")
json.dump(tj, open(tj_path, "w"), indent=1)
print("wrote", tj_path)
