"""Condense an `ncu --set full` report (or a `--metrics gpu__time_duration.sum` launch list) into the
small text files kept under profiles/.

    python tools/ncu_summary.py full  gpurun_out/x.ncu-rep  profiles/rNN_x.md   [note ...]
    python tools/ncu_summary.py list  gpurun_out/launches.csv profiles/rNN_launches.md [note ...]
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit (regs), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occ limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (DMMA) active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor pipe (HMMA/TF32) inst %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe inst %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__sass_inst_executed_op_utcmma.sum", "tcgen05.mma instructions (UTCxMMA)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor memory (TMEM) active %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor-core shared-memory operand wavefronts % of peak"),
    ("sm__ops_path_tensor_op_hmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "mma.sync TF32 ops % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]


def _raw_csv(path):
    if path.endswith(".csv"):   # a raw page already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.raw.csv
        rows = list(csv.reader(open(path)))
        return rows[0], rows[1], rows[2:]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full(path, out, note):
    hdr, units, rows = _raw_csv(path)
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{path.split('/')[-1]}`", ""]
    if note:
        lines += [note, ""]
    lines += ["Numbers under ncu are cold-cache / serialised replays: use them for shares and pipe utilisation, "
              "never as bench values.", ""]
    for r in rows:
        name = r[col["Kernel Name"]]
        lines += [f"## `{name}`", "", "| metric | value |", "|---|---|"]
        for key, label in FULL_KEYS:
            if key in col and r[col[key]] != "":
                lines.append(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
        if "dram__bytes_read.sum" in col:
            def gb(k):
                v, u = float(r[col[k]]), units[col[k]].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            lines.append(f"| **DRAM traffic per launch (read+write)** | {(gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')) / 1e9:.4f} GB |")
        lines.append("")
    open(out, "w").write("\n".join(lines))


def launch_list(path, out, note):
    rows = [r for r in csv.reader(open(path)) if r]
    # find the header row
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        u = r[col["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = r[col["Kernel Name"]].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    lines = [f"# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of `{path.split('/')[-1]}`", ""]
    if note:
        lines += [note, ""]
    lines += ["Per-launch times under ncu are cold-cache and serialised: the SHARE per kernel is what must agree "
              "with bench.py's CUDA-event shares, not the absolute.", "",
              "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {ms:.3f} | {ms / n:.4f} | {ms / total:.4f} |")
    lines += ["", f"total {total:.3f} ms over {sum(a[0] for a in agg.values())} launches", ""]
    open(out, "w").write("\n".join(lines))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = " ".join(sys.argv[4:])
    {"full": full, "list": launch_list}[mode](src, dst, note)
    print("wrote", dst)
