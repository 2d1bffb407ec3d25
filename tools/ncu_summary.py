"""Dump the metrics we quote from an .ncu-rep into a small text file (profiles/ is what gets committed)."""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
    'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def main(rep, out=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"== {d.get('Kernel Name', '?')[:110]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in hdr:
            if k in KEYS or k.startswith('sm__mem_tensor_cycles_active.avg') or 'pipe_tensor' in k and '.avg.pct' in k or ('warp_issue_stalled' in k and k.endswith('_per_warp_active.pct')):
                try:
                    v = float(d[k])
                except Exception:
                    continue
                if 'stalled' in k and v < 2.0:
                    continue
                lines.append(f'   {k:78s} {v:18.3f} {u.get(k, "")}')
    text = '\n'.join(lines) + '\n'
    if out:
        open(out, 'w').write(text)
    else:
        print(text)


if __name__ == '__main__':
    main(*sys.argv[1:3])
