"""Print the handful of ncu metrics we track from a --page raw --csv export (one row per kernel)."""
import csv, json, os, re, sys
# usage: ncu_summary.py raw.csv [traffic.json]   (the second argument accumulates {kernel: dram bytes per launch})
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_uniform.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for vals in rows[2:]:
    print('-' * 100)
    for i, h in enumerate(hdr):
        if h in keys or h.startswith('l1tex__data_pipe_lsu_wavefronts') or h.startswith('l1tex__data_bank_conflicts') \
                or h.startswith('l1tex__m_xbar2l1tex_read_bytes') or h.startswith('lts__t_sectors_srcunit_tex_op_read.sum') \
                or (h.startswith('smsp__warp_issue_stalled') and h.endswith('per_warp_active.pct')) \
                or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try:
                v = float(vals[i].replace(',', ''))
                if h.startswith('smsp__warp_issue_stalled') and v < 2.0:
                    continue
            except ValueError:
                pass
            print(f'{h:85s} {vals[i]:>20s} {units[i]}')

if len(sys.argv) > 2:
    path = sys.argv[2]
    tr = json.load(open(path)) if os.path.exists(path) else {}
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tmul = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    col = {h: i for i, h in enumerate(hdr)}
    for vals in rows[2:]:
        m = re.search(r'([A-Za-z_0-9]+)\s*(<|\()', vals[col['Kernel Name']].replace('void ', '').replace('impdar::', ''))
        if not m:
            continue
        def val(h, table):
            return float(vals[col[h]].replace(',', '')) * table.get(units[col[h]], 1.0)
        pipes = {}
        for h in hdr:
            if h in ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
                     'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
                     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
                     'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
                     'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
                     'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum'):
                try:
                    pipes[h] = float(vals[col[h]].replace(',', ''))
                except ValueError:
                    pass
        tr[m.group(1)] = {'dram_bytes': val('dram__bytes_read.sum', mult) + val('dram__bytes_write.sum', mult),
                          'ncu_ms': val('gpu__time_duration.sum', tmul), 'source': os.path.basename(sys.argv[1]),
                          'pipes': pipes}
    json.dump(tr, open(path, 'w'), indent=1, sort_keys=True)
