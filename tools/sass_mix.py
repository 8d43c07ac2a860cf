#!/usr/bin/env python3
"""Static instruction mix of one kernel, by source region (no GPU needed).

    python tools/sass_mix.py [--obj sushi_b200/csrc/sb_fused2.o] [--kernel k_match_packedIhE] [--top 12]

Disassembles the object with `nvdisasm -gi` (the objects are built with -lineinfo) and attributes every SASS
instruction to a named region of sb_fused2.cu through its chain of inlining locations: the region is the first
entry of the chain (innermost first) that falls inside one of REGIONS, so helper functions (fma2, cmul_s, ...)
count towards the phase that called them.  The counts are STATIC (one per instruction in the binary): loops
that stay rolled (`#pragma unroll 1`: the multiply loop over partitions) count once, everything in the FFT
passes and the epilogue is fully unrolled, so for those phases static counts per thread are dynamic counts
per thread.  Used to check instruction-level changes of the issue-bound phases before spending GPU time.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (name, first line, last line) in sb_fused2.cu; first match wins, innermost location first
REGIONS = [
    ('epilogue 3: windows, totals, scan, bases', 'finish_item_v3', 'mbar_wait(s_bar, bar_parity);\n    // ---- the 32 + 32 window bytes', 'float run_lb[RUNS];'),
    ('epilogue 3: per run radix-2 step + bounds', 'finish_item_v3', 'float run_lb[RUNS];', '// ---- block minimum of the upper bounds'),
    ('epilogue 3: block minimum, selection', 'finish_item_v3', '// ---- block minimum of the upper bounds', '// the selected runs leave as records'),
    ('epilogue 3: records', 'finish_item_v3', '// the selected runs leave as records', 'unsigned long long best = ~0ull;\n    if (__any_sync'),
    ('epilogue 3: in-kernel path for what could not leave', 'finish_item_v3', 'unsigned long long best = ~0ull;\n    if (__any_sync', '// ---------------------------------------------------------------- kernel A:'),
    ('epilogue 3: setup', 'finish_item_v3', '__device__ __forceinline__ void finish_item_v3', 'mbar_wait(s_bar, bar_parity);\n    // ---- the 32 + 32 window bytes'),
    ('epilogue 1 (first version)', 'finish_item', '__device__ __forceinline__ void finish_item(', '// Body 3 (uint8 streams).  A thread owns'),
    ('fft passes', 'fft_passes_dif', '__device__ __forceinline__ void fft_passes_dif', '// Body 3: the constants of a query every thread needs in the epilogue'),
    ('stage inputs', 'stage_inputs', '__device__ __forceinline__ void stage_inputs', '// Y += conj(T) * X on both slots'),
]


def resolve_regions(src_path):
    lines = open(src_path).read().split('\n')

    text = '\n'.join(lines)

    def find(needle, start=0):
        """1-based line of the first occurrence of `needle` (may span lines) at or after line index `start`."""
        pos = text.find(needle, sum(len(l) + 1 for l in lines[:start]))
        if pos < 0:
            raise SystemExit('marker not found in %s: %r' % (src_path, needle))
        return text.count('\n', 0, pos) + 1
    out = []
    for name, _, a, b in REGIONS:
        la = find(a)
        lb = find(b, la) - 1
        out.append((name, la, lb))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--obj', default=os.path.join(ROOT, 'sushi_b200', 'csrc', 'sb_fused2.o'))
    ap.add_argument('--src', default=os.path.join(ROOT, 'sushi_b200', 'csrc', 'sb_fused2.cu'))
    ap.add_argument('--kernel', default='k_match_pairIhLi3E', help='substring of the mangled kernel name')
    ap.add_argument('--top', type=int, default=10, help='opcodes listed per region')
    args = ap.parse_args()

    regions = resolve_regions(args.src)
    src_name = os.path.basename(args.src)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(args.obj)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = [f for f in os.listdir(tmp) if f.endswith('.cubin')]
        text = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubins[0])], stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, check=True).stdout.decode()

    in_fn = False
    chain = []          # current location chain, innermost first
    fresh = True        # the next File line starts a new chain
    loc_re = re.compile(r'File "([^"]+)", line (\d+)')
    ins_re = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)')
    by_region = collections.OrderedDict((r[0], collections.Counter()) for r in regions)
    by_region['other (multiply, packing, kernel body)'] = collections.Counter()
    total = 0
    for line in text.split('\n'):
        if line.startswith('.text.'):
            in_fn = args.kernel in line
            chain, fresh = [], True
            continue
        if not in_fn:
            continue
        if '//## File' in line:
            if fresh:
                chain, fresh = [], False
            for f, n in loc_re.findall(line):
                chain.append((os.path.basename(f), int(n)))
            continue
        m = ins_re.match(line)
        if not m:
            continue
        fresh = True
        op = m.group(1)
        region = 'other (multiply, packing, kernel body)'
        for f, n in chain:
            if f != src_name:
                continue
            hit = next((r[0] for r in regions if r[1] <= n <= r[2]), None)
            if hit:
                region = hit
                break
        by_region[region][op.split('.')[0]] += 1
        total += 1
    if not total:
        raise SystemExit('kernel not found: ' + args.kernel)
    print('%s: %d SASS instructions' % (args.kernel, total))
    for name, c in by_region.items():
        n = sum(c.values())
        if not n:
            continue
        print('  %-40s %6d  %s' % (name, n, ' '.join('%s:%d' % kv for kv in c.most_common(args.top))))


if __name__ == '__main__':
    sys.exit(main())
