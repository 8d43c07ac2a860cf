#!/usr/bin/env python3
"""Turn an `ncu --set full` report of the dominant kernel, captured on the bench workload, into
profiles/traffic.json (DRAM bytes per launch) + a short text summary under profiles/.
    ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 \
        -o gpurun_out/pair_bench python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-load-leg    (GPU box)
    python tools/ncu_traffic.py gpurun_out/pair_bench.ncu-rep config3 uint8 match_fused 1 [src_hash [pairs_in_launch]]   (here)
The entry is stamped with the hash of the CUDA sources (bench.kernel_source_hash): bench.py uses a capture only
for the source it was taken from and prints traffic = null with the reason otherwise.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload, stype, kclass, ngpus = sys.argv[1:6]
src_hash = sys.argv[6] if len(sys.argv) > 6 else None      # hash of the CUDA sources the capture was taken from (gpu_pass.sh writes it)
n_units = float(sys.argv[7]) if len(sys.argv) > 7 else None  # pairs of lag blocks the captured launch processed (the kernel is persistent: its grid is the SM count)
sys.path.insert(0, ROOT)
import bench  # noqa: E402
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name):
    v, u = d[name]
    v = float(v)
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9,
             's': 1.0, 'second': 1.0, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9}.get(u, 1.0)
    return v * scale


dram = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
out_path = os.path.join(ROOT, 'profiles', 'traffic.json')
try:
    tr = json.load(open(out_path))
except (OSError, ValueError):
    tr = {}
key = '%s/%s/%s/N%s' % (workload, stype, kclass, ngpus)
keep = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
grid = float(d['launch__grid_size'][0])
n_units = n_units or grid
tr[key] = {'dram_bytes_per_launch': dram, 'grid_size': grid, 'pairs_in_launch': n_units, 'dram_bytes_per_pair': dram / n_units, 'report': os.path.basename(rep), 'src_hash': src_hash or bench.kernel_source_hash(),
           'metrics': {k: ' '.join(d[k]) for k in keep if k in d}}
json.dump(tr, open(out_path, 'w'), indent=1, sort_keys=True)
print(key, 'dram bytes/launch', dram)
for k in keep:
    if k in d:
        print('  %-70s %s %s' % (k, d[k][0], d[k][1]))
