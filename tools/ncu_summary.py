#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into a small text table
for profiles/.   python tools/ncu_summary.py gpurun_out/<tag>/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print('# source: %s (ncu --set full --clock-control none; per-launch, cold-ish cache, serialised)' % path)
    for d in data:
        print('\n== %s' % d[hdr.index('Kernel Name')][:110])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print('%-78s %16s %s' % (m, d[i], units[i]))
        rd = float(d[hdr.index('dram__bytes_read.sum')])
        wr = float(d[hdr.index('dram__bytes_write.sum')])
        u = units[hdr.index('dram__bytes_read.sum')]
        print('%-78s %16.3f %s' % ('traffic = dram read + write', rd + wr, u))


def traffic(path, out_json, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by kernel
    family, merged into `out_json` (what bench.py reports as roofline.traffic)."""
    import json
    import os
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    db = json.load(open(out_json)) if os.path.exists(out_json) else {}
    entry = db.setdefault(workload, {})
    for d in data:
        name = d[hdr.index('Kernel Name')]
        key = 'msda_bwd_rows_kernel' if 'msda_bwd' in name else 'msda_fwd_rows_kernel'
        tot = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(m)
            tot += float(d[i]) * scale[units[i]]
        entry[key] = tot
    entry['source'] = os.path.basename(path)
    json.dump(db, open(out_json, 'w'), indent=1, sort_keys=True)
    print(json.dumps(db[workload]))


if __name__ == '__main__':
    if len(sys.argv) >= 5 and sys.argv[1] == '--traffic':
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1])
