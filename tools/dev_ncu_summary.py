"""Developer helper (not a test): summarise an .ncu-rep (raw + source pages) the way profiles/*.txt are written.
   python tools/dev_ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, 'raw')
    h, u, v = rows[0], rows[1], rows[2]
    d = dict(zip(h, zip(u, v)))
    print('kernel:', d['Kernel Name'][1])
    for k in KEYS:
        if k in d:
            print('%-75s %-12s %s' % (k, d[k][0], d[k][1]))
    for k in h:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
            val = float(d[k][1])
            if val > 0.05:
                print('%-75s %.3f' % (k.replace('smsp__average_warps_issue_stalled_', 'stall/issue: ').replace('_per_issue_active.ratio', ''), val))
    src = page(rep, 'source')
    hh, data = src[1], src[2:]
    ie, isamp, isrc = hh.index('Instructions Executed'), hh.index('# Samples'), hh.index('Source')
    tot = sum(int(r[ie]) for r in data)
    tsamp = sum(int(r[isamp]) for r in data)
    print('SASS instructions: %d, executed warp-instructions %.4e, samples %d' % (len(data), tot, tsamp))
    blocks, cur = [], None
    for i, r in enumerate(data):
        e, s = int(r[ie]), int(r[isamp])
        if cur and abs(e - cur['e']) <= 0.02 * max(e, cur['e'], 1):
            cur['n'] += 1; cur['sum'] += e; cur['samp'] += s; cur['end'] = i
        else:
            cur = {'start': i, 'end': i, 'e': e, 'n': 1, 'sum': e, 'samp': s}; blocks.append(cur)
    print('blocks of SASS with equal execution count (>0.5% of instructions or >1% of samples):')
    for b in blocks:
        if b['sum'] > 0.005 * tot or b['samp'] > 0.01 * tsamp:
            print('  sass %5d-%5d  n=%4d  exec/inst=%.3e  inst share=%5.1f%%  sample share=%5.1f%%' % (
                b['start'], b['end'], b['n'], b['e'], 100.0 * b['sum'] / tot, 100.0 * b['samp'] / tsamp))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:12]
    print('hottest SASS by samples:')
    for i in top:
        print('  %5d  %6.2f%%  %s' % (i, 100.0 * int(data[i][isamp]) / tsamp, data[i][isrc][:100]))


if __name__ == '__main__':
    main()
