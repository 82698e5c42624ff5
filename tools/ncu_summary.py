"""Summarise `ncu --set full` reports (.ncu-rep) into a small CSV of raw metrics per kernel launch, and record the DRAM bytes per
launch that bench.py quotes as `roofline.traffic`.

    python tools/ncu_summary.py profiles/ncu_r02_kernels.csv profiles/r02_kernel_traffic.json  rep1.ncu-rep [rep2.ncu-rep ...]
"""
import csv
import json
import os
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'time_ns'),
    ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
    ('dram__bytes_read.sum', 'dram_read_bytes'),
    ('dram__bytes_write.sum', 'dram_write_bytes'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct_active'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pct_elapsed'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_bytes'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct'),
    ('smsp__inst_executed.sum', 'warp_insts'),
]
UNIT = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1, 'us': 1e3, 'ms': 1e6, 'ns': 1, 'usecond': 1e3, 'nsecond': 1, 'msecond': 1e6}


def rows_of(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    r = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = r[0], r[1]
    col = {h: i for i, h in enumerate(hdr)}
    for row in r[2:]:
        d = {'kernel': row[col['Kernel Name']].split('(')[0], 'report': os.path.basename(rep)}
        for m, name in WANT:
            if m in col and row[col[m]] != '':
                v = float(row[col[m]].replace(',', ''))
                d[name] = v * UNIT.get(units[col[m]], 1) if name.endswith(('bytes', '_ns')) else v
        yield d


def main():
    out_csv, out_json, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    rows = [d for rep in reps for d in rows_of(rep)]
    names = ['kernel', 'report'] + [n for _, n in WANT]
    with open(out_csv, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=names)
        w.writeheader()
        for d in rows:
            w.writerow({k: d.get(k, '') for k in names})
    traffic = json.load(open(out_json)) if os.path.exists(out_json) else {}
    for d in rows:
        key = d['kernel'].replace('void ', '').replace('scf::', '').split('<')[0]
        traffic[key] = {'dram_read_bytes': d.get('dram_read_bytes'), 'dram_write_bytes': d.get('dram_write_bytes'),
                        'time_ns_under_ncu': d.get('time_ns'), 'tensor_pct_active': d.get('tensor_pct_active'),
                        'source': f'{os.path.basename(out_csv)} <- ncu --set full, {d["report"]} (per launch, B=32)'}
    with open(out_json, 'w') as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    for d in rows:
        print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.items()})


if __name__ == '__main__':
    main()
