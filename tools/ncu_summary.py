"""Summarise an .ncu-rep (ncu --set full --import-source on) into JSON: launch facts, pipe
utilisation, DRAM traffic and the dynamic SASS opcode histogram per unit of work.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --units 1e8 > profiles/x.json
(--units = run-steps (or samples) one launch processes; histogram is printed per unit.)"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
       'launch__block_size', 'launch__occupancy_limit_registers',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
       'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
       'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
       'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
       'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def ncu(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(out.stdout)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[sys.argv.index('--units') + 1]) if '--units' in sys.argv else None
    launch = int(sys.argv[sys.argv.index('--launch') + 1]) if '--launch' in sys.argv else 0
    rows = ncu(rep, 'raw')
    hdr, unit_row = rows[0], rows[1]
    r = rows[2 + launch]
    res = {'report': rep, 'kernel': r[hdr.index('Kernel Name')], 'launches_in_report': len(rows) - 2}
    for k in RAW:
        if k in hdr:
            res[k] = '%s %s' % (r[hdr.index(k)], unit_row[hdr.index(k)])
    src = ncu(rep, 'source')
    # the source page lists one block per launch; take the first
    h = src[1]
    ia, ie = h.index('Source'), h.index('Instructions Executed')
    by, tot = collections.Counter(), 0
    for row in src[2:]:
        if len(row) <= ie or row[0] == 'Kernel Name':
            if tot:
                break
            continue
        try:
            n = int(row[ie])
        except ValueError:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', row[ia].strip())
        op = (m.group(2) if m else row[ia][:12]).split('.')[0]
        by[op] += n
        tot += n
    res['warp_instructions'] = tot
    if units:
        res['units_per_launch'] = units
        res['thread_instructions_per_unit'] = tot * 32 / units
        fp64 = sum(v for k, v in by.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
        res['fp64_thread_instructions_per_unit'] = fp64 * 32 / units
        res['opcode_thread_instructions_per_unit'] = {k: round(v * 32 / units, 2)
                                                      for k, v in by.most_common(24)}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
