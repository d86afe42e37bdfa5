#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into a JSON list + a markdown table for profiles/."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'time_us',
    'dram__bytes_read.sum': 'dram_bytes_read',
    'dram__bytes_write.sum': 'dram_bytes_write',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'dram__bytes_read.sum.per_second': 'dram_read_per_s',
    'dram__bytes_write.sum.per_second': 'dram_write_per_s',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_pct_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__registers_per_thread': 'regs',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem',
    'lts__t_bytes.sum': 'l2_bytes',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts',
    'smsp__inst_executed.sum': 'warp_insts',
}
SCALE = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1, 'Tbyte': 1e12, 'Kbyte/s': 1e3, 'Mbyte/s': 1e6, 'Gbyte/s': 1e9,
         'Tbyte/s': 1e12, 'byte/s': 1, 'us': 1, 'ns': 1e-3, 'ms': 1e3, 'Kbyte/block': 1e3, 'byte/block': 1, 'Mbyte/block': 1e6}


def summarise(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {'report': path.split('/')[-1], 'kernel': vals[hdr.index('Kernel Name')][:60]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                try:
                    d[KEYS[h]] = float(v.replace(',', '')) * SCALE.get(u, 1)
                except ValueError:
                    d[KEYS[h]] = v
        res.append(d)
    return res


if __name__ == '__main__':
    allr = []
    for p in sys.argv[1:]:
        allr += summarise(p)
    print(json.dumps(allr, indent=1))
