"""summarise ncu CSV exports (raw + source pages) into markdown: key metrics per kernel + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/<tag> [cu-file-stem for line mapping]"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']


def I(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def line_map(cu_stem, kernel_substr):
    """SASS instruction order -> source line, from nvdisasm --print-line-info of the in-tree cubin"""
    tmp = Path('/tmp/sass_map')
    tmp.mkdir(exist_ok=True)
    subprocess.run(f'cd {tmp} && rm -f *.cubin && cuobjdump -xelf all /root/repo/vren_b200/libvrenb200.so > /dev/null', shell=True)
    cub = tmp / f'{cu_stem}.sm_100a.cubin'
    if not cub.exists():
        return None
    out = subprocess.run(['nvdisasm', '--print-line-info', '-c', str(cub)], capture_output=True, text=True).stdout.split('\n')
    start = None
    for i, l in enumerate(out):
        if l.startswith('.text.') and kernel_substr in l:
            start = i
            break
    if start is None:
        return None
    seq, cur = [], None
    for l in out[start + 1:]:
        if l.startswith('.section') or l.startswith('.text.'):
            break
        m = re.search(r'//## File ".*?", line (\d+)', l)
        if m:
            cur = int(m.group(1))
            continue
        if re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l):
            seq.append(cur)
    return seq


def main():
    tag = sys.argv[1]
    raw = list(csv.reader(open(tag + '_raw.csv')))
    hdr, units = raw[0], raw[1]
    print(f'# ncu summary: {tag}\n')
    for r in raw[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f'## {name[:110]}\n\n| metric | value | unit |\n|---|---|---|')
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f'| {k} | {r[i]} | {units[i]} |')
        print()
    src = list(csv.reader(open(tag + '_src.csv')))
    # split per kernel: rows starting with "Kernel Name"
    blocks, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    stems = {'exclusive_scan': 'scan', 'find_unique': 'clustered', 'assign_lights': 'clustered', 'onesweep': 'radix_sort',
             'radix_histogram': 'radix_sort', 'depth_pyramid': 'depth_pyramid', 'build_bvh': 'bvh', 'reduce_': 'reduce'}
    for b in blocks:
        rows = b['rows']
        if not rows:
            continue
        h = rows[0]
        data = [r for r in rows[1:] if len(r) == len(h)]
        ci = {x: i for i, x in enumerate(h)}
        stall = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
        tot_i = sum(I(r[ci['Instructions Executed']]) for r in data) or 1
        tot_s = sum(I(r[ci['# Samples']]) for r in data) or 1
        print(f"## source view: {b['name'][:100]}\n\ninstructions {tot_i}, samples {tot_s}\n")
        agg = collections.Counter()
        for r in data:
            for c in stall:
                agg[c] += I(r[ci[c]])
        print('stall mix: ' + ', '.join(f'{k[6:]} {v / tot_s * 100:.0f}%' for k, v in agg.most_common(7)) + '\n')
        seq = None
        # demangled "ns::<unnamed>::kernel<(int)256, (int)24, ...>(" -> mangled fragment "kernelILi256ELi24E..."
        m = re.search(r'(\w+_kernel)(?:<([^>]*)>)?\(', b['name'])
        if m:
            sub = m.group(1)
            if m.group(2):
                args = re.findall(r'\((?:int|bool)\)(\d+)', m.group(2))
                bools = re.findall(r'\(bool\)', m.group(2))
                if args and not bools:
                    sub += 'I' + ''.join(f'Li{a}E' for a in args)
            for key, stem in stems.items():
                if key in b['name']:
                    seq = line_map(stem, sub)
                    cu = Path(f'/root/repo/vren_b200/csrc/{stem}.cu').read_text().split('\n')
                    break
        if seq and len(seq) == len(data):
            per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
            for ln, r in zip(seq, data):
                per[ln][0] += I(r[ci['Instructions Executed']])
                per[ln][1] += I(r[ci['# Samples']])
                for c in stall:
                    per[ln][2][c] += I(r[ci[c]])
            print('| line | inst % | samples % | top stall | source |\n|---|---|---|---|---|')
            for ln, (a, s, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[:22]:
                top = st.most_common(1)[0][0][6:] if st else ''
                text = cu[ln - 1].strip()[:80].replace('|', '/') if ln and ln <= len(cu) else ''
                print(f'| {ln} | {a / tot_i * 100:.1f} | {s / tot_s * 100:.1f} | {top} | `{text}` |')
        else:
            top = sorted(data, key=lambda r: -I(r[ci['# Samples']]))[:15]
            print('| samples % | inst | SASS |\n|---|---|---|')
            for r in top:
                print(f"| {I(r[ci['# Samples']]) / tot_s * 100:.1f} | {r[ci['Instructions Executed']]} | `{r[ci['Source']][:70]}` |")
        print()


if __name__ == '__main__':
    main()
