#!/usr/bin/env python3
"""Compact per-kernel summary of an .ncu-rep (read on the CPU box), units normalised (us, MB):

    python tools/ncu_summary.py rep.ncu-rep [out.txt [traffic.json workload]]

With `traffic.json workload` the DRAM bytes (read + write) per launch of every kernel in the capture are merged into
profiles/traffic.json under that bench.py workload name (bench.py copies them into `roofline.traffic`)."""
import csv
import json
import re
import subprocess
import sys

WANT = {
    'Kernel Name': 'name', 'gpu__time_duration.sum': 'us', 'launch__registers_per_thread': 'regs',
    'launch__grid_size': 'grid', 'launch__block_size': 'blk',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'dram__bytes_read.sum': 'dramR_MB',
    'dram__bytes_write.sum': 'dramW_MB', 'smsp__inst_executed.sum': 'inst',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor%',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'st_long',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'st_short',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'st_bar',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'st_wait',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'st_mio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'st_math',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'st_notsel',
    'l1tex__t_sector_hit_rate.pct': 'l1hit', 'lts__t_sector_hit_rate.pct': 'l2hit',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'bankconf',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wf',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram%',
}
TIME = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 's': 1e6, 'second': 1e6}
BYTES = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'Tbyte': 1e6}


def num(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {v: hdr.index(k) for k, v in WANT.items() if k in hdr}
    lines, traffic = [], {}
    for r in rows[2:]:
        d = {}
        for k, i in idx.items():
            v = r[i]
            if k == 'us' and num(v) is not None:
                v = num(v) * TIME.get(units[i], 1.0)
            elif k in ('dramR_MB', 'dramW_MB') and num(v) is not None:
                v = num(v) * BYTES.get(units[i], 1.0)
            d[k] = v
        name = d.pop('name')
        fmt = lambda x: ('%.4g' % x) if isinstance(x, float) else (('%.3g' % num(x)) if num(x) is not None else x)
        lines.append(name.replace('void ', '').replace('nnue::', '')[:52] + ' | ' + ' '.join('%s=%s' % (k, fmt(v)) for k, v in d.items()))
        m = re.search(r'(\w+)\s*(<|\()', name.replace('void ', '').replace('nnue::', ''))
        if m and isinstance(d.get('dramR_MB'), float):
            traffic.setdefault(m.group(1), []).append((d['dramR_MB'] + d['dramW_MB']) * 1e6)
    out = '\n'.join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(out + '\n')
    if len(sys.argv) > 4:
        path, workload = sys.argv[3], sys.argv[4]
        try:
            old = json.load(open(path))
        except Exception:
            old = {}
        cur = old.setdefault(workload, {})
        for k, v in traffic.items():
            cur[k] = max(v)  # several launches of one kernel in a step (different operands): the largest is the named stage's
        json.dump(old, open(path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
