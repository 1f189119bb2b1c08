#!/usr/bin/env python3
"""Compact per-kernel summary of an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep [out.txt]"""
import csv
import subprocess
import sys

WANT = {
    'Kernel Name': 'name', 'gpu__time_duration.sum': 'us', 'launch__registers_per_thread': 'regs',
    'launch__grid_size': 'grid', 'launch__block_size': 'blk',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'dram__bytes_read.sum': 'dramR_MB',
    'dram__bytes_write.sum': 'dramW_MB', 'smsp__inst_executed.sum': 'inst',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'st_long',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'st_short',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'st_bar',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'st_wait',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'st_mio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'st_math',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'st_lg',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'st_notsel',
    'l1tex__t_sector_hit_rate.pct': 'l1hit', 'lts__t_sector_hit_rate.pct': 'l2hit',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'bankconf',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wf',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram%',
    'l1tex__data_pipe_lsu_wavefronts.sum': 'lsu_wf',
    'sm__cycles_elapsed.avg': 'cycles',
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {v: hdr.index(k) for k, v in WANT.items() if k in hdr}
    lines = []
    for r in rows[2:]:
        d = {k: r[i] for k, i in idx.items()}
        name = d.pop('name')[:44]

        def f(x):
            try:
                return '%.3g' % float(x.replace(',', ''))
            except Exception:
                return x
        lines.append(name + ' | ' + ' '.join('%s=%s' % (k, f(v)) for k, v in d.items()))
    out = '\n'.join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(out + '\n')
    if len(sys.argv) > 3:  # traffic.json: kernel function name -> dram bytes (read + write) per launch
        import json
        import re
        traffic = {}
        ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        unit_r, unit_w = rows[1][ir], rows[1][iw]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        for r in rows[2:]:
            m = re.search(r'(\w+)\s*(<|\()', r[hdr.index('Kernel Name')].replace('void ', '').replace('nnue::', ''))
            if m:
                traffic[m.group(1)] = float(r[ir].replace(',', '')) * scale.get(unit_r, 1.0) + \
                    float(r[iw].replace(',', '')) * scale.get(unit_w, 1.0)
        old = {}
        try:
            old = json.load(open(sys.argv[3]))
        except Exception:
            pass
        old.update(traffic)
        json.dump(old, open(sys.argv[3], 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
