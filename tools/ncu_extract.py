#!/usr/bin/env python
"""Raw-page extract of an ncu report as a small JSON (what profiles/ keeps; the .ncu-rep files stay in gpurun_out/):
    python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/x_ncu_full.json [launch index] [extra metric substrings ...]
Keeps the headline metrics (duration, DRAM bytes, L2 / shared-memory / tensor pipes, issue and occupancy figures, stall samples)."""
import csv, io, json, subprocess, sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_red.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active')


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].lstrip('-').isdigit() else 0
    extra = [a for a in sys.argv[3:] if not a.lstrip('-').isdigit()]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    r = data[which]
    m = {}
    for i, n in enumerate(head):
        if n in KEEP or n.startswith('smsp__pcsamp_warps_issue_stalled') and not n.endswith('not_issued') or any(e in n for e in extra):
            if r[i] != '':
                m[n] = {'value': r[i], 'unit': units[i]}
    json.dump({'kernel': r[head.index('Kernel Name')], 'source': rep, 'launch': which, 'launches_in_report': len(data), 'metrics': m},
              open(out, 'w'), indent=1)
    print(out, len(m), 'metrics;', r[head.index('Kernel Name')][:80])


if __name__ == '__main__':
    main()
