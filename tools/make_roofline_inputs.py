"""profiles/roofline_inputs_r02.json from an ncu summary (tools/ncu_summary.py output) of the dominant
kernel at bench.py's workload: the two numbers bench.py's roofline fields need, with the launch shape
they are valid for and the file they come from.
    python tools/make_roofline_inputs.py profiles/ncu_mc_spec_r02_cfg2_g4_p6.json 4 6,1,0 [profiles/ncu_r02_mc_plain.json]
The optional last file is the summary of the one-lane-per-run kernel: its FP64 count per run-step is the
work that is not replicated across a lane group (bench.py's roofline_fp64.frac_nonreplicated)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, lanes, shape = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    d = json.load(open(src))
    num = lambda k: float(d[k].split()[0]) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6}[d[k].split()[1]]  # noqa: E731
    out = {'lanes_per_run': lanes, 'shape': shape, 'kernel': d['kernel'],
           'fp64_thread_instructions_per_run_step': d['fp64_thread_instructions_per_unit'],
           'thread_instructions_per_run_step': d['thread_instructions_per_unit'],
           'dram_bytes_per_launch': num('dram__bytes_read.sum') + num('dram__bytes_write.sum'),
           'source': os.path.relpath(src, ROOT) + ' (ncu --set full --import-source on, one launch of %s)' % d['kernel']}
    if len(sys.argv) > 4:
        g1 = json.load(open(sys.argv[4]))
        out['fp64_thread_instructions_per_run_step_one_lane'] = g1['fp64_thread_instructions_per_unit']
        out['one_lane_source'] = os.path.relpath(sys.argv[4], ROOT) + ' (%s)' % g1['kernel']
    with open(os.path.join(ROOT, 'profiles', 'roofline_inputs_r02.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
