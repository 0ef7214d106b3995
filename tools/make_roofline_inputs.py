"""profiles/roofline_inputs_r02.json from an ncu summary (tools/ncu_summary.py output) of the dominant
kernel at bench.py's workload: the two numbers bench.py's roofline fields need, with the launch shape
they are valid for and the file they come from.
    python tools/make_roofline_inputs.py profiles/ncu_mc_spec_r02_cfg2_g4_p6.json 4 6,1,0"""
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
    with open(os.path.join(ROOT, 'profiles', 'roofline_inputs_r02.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
