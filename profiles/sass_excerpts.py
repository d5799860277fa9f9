'''
SASS evidence for the kernels of the day: mnemonic counts (memory / synchronisation / fp64) and a short excerpt of each kernel's
main loop, from `cuobjdump -sass` of the built library.   python profiles/sass_excerpts.py > profiles/r2/sass_excerpts.txt
'''
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'covasim_b200', 'libcovasim_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=tmp, capture_output=True)
out = []


def summarize(title, cubin, pattern, excerpt_re=None, n_excerpt=40):
    path = os.path.join(tmp, cubin)
    syms = subprocess.run(['cuobjdump', '-elf', path], capture_output=True, text=True).stdout
    name = sorted(set(re.findall(r'\.text\.(\S*' + pattern + r'\S*)', syms)))[0]
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', name, path], capture_output=True, text=True).stdout
    ins = [l for l in txt.splitlines() if re.match(r'\s+/\*[0-9a-f]{4}\*/', l)]
    mn = collections.Counter()
    for l in ins:
        m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            mn[m.group(1)] += 1
    prefixes = ('LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'CCTL', 'LDGSTS', 'UBLKCP', 'UTMA', 'SYNCS', 'VOTE', 'SHFL', 'BAR', 'DMUL', 'DFMA', 'DADD', 'MUFU',
                'IMAD.WIDE', 'ACQBULK', 'LDGDEPBAR', 'PREFETCH')
    keys = [k for k in mn if k.startswith(prefixes)]
    tma = [k for k in mn if k.startswith(('UBLKCP', 'UTMA', 'LDGSTS', 'SYNCS'))]
    out.extend(['=' * 110, title, 'function: ' + name, 'SASS instructions: %d' % len(ins),
                'memory / sync / fp64 mnemonics: ' + ', '.join('%s x%d' % (k, mn[k]) for k in sorted(keys)),
                'TMA / bulk-copy / mbarrier instructions: ' + (', '.join(tma) if tma else 'none')])
    if excerpt_re:
        idx = [i for i, l in enumerate(ins) if re.search(excerpt_re, l)]
        if idx:
            a = max(idx[0] - 4, 0)
            out.append('excerpt (around the first match of %s):' % excerpt_re)
            out.extend(l.rstrip()[:150] for l in ins[a:a + n_excerpt])
    out.append('')


summarize('dense streaming edge pass (dynamic layers / use_adjacency=False): edge_pass_kernel<MULTI=false, SMEM_BITS=true, 512 threads, 2 quads, L2 prefetch 1 tile>',
          'edge_pass.sm_100a.cubin', 'edge_pass_kernelILb0ELb1ELi512ELi2ELi1', excerpt_re=r'LDG\.E\.128', n_excerpt=46)
summarize('sparse edge pass of the fused day (adjacency form, transmitter entries): edge_pass_sparse2_kernel<MULTI=false, 16 lanes, 4 entries in flight>',
          'edge_pass.sm_100a.cubin', 'edge_pass_sparse2_kernelILb0ELi16ELi4', excerpt_re=r'LDG\.E\.128', n_excerpt=40)
summarize('day_begin_kernel<END, PRE, TEST, TSEL> (first per-agent kernel of the fused day)', 'day_fused.sm_100a.cubin', 'day_begin_kernelILb1ELb1ELb1ELb1',
          excerpt_re=r'ACQBULK|LDG\.E ', n_excerpt=30)
summarize('day_mid_kernel (second per-agent kernel of the fused day)', 'day_fused.sm_100a.cubin', 'day_mid_kernel', excerpt_re=r'ACQBULK|LDG\.E ', n_excerpt=24)
summarize('dense edge pass with bulk-copy staging (measured slower, CVB_DENSE_VARIANT=3): edge_pass_tma_kernel<MULTI=false, 20 warps>',
          'edge_pass.sm_100a.cubin', 'edge_pass_tma_kernelILb0ELi20', excerpt_re=r'UBLKCP', n_excerpt=24)
summarize('agent-partitioned edge pass: edge_pass_partition_kernel<MULTI=true, 8 lanes, 2 entries in flight>', 'edge_pass.sm_100a.cubin',
          'edge_pass_partition_kernelILb1ELi8ELi2', excerpt_re=r'LDG\.E\.128', n_excerpt=30)
summarize('exchange over peer memory: peer_push_kernel<uint4> (16-byte stores into every rank\'s buffer)', 'capi.sm_100a.cubin', 'peer_push_kernelI5uint4',
          excerpt_re=r'STG\.E\.128', n_excerpt=16)
hdr = ['SASS evidence for the round-2 kernels (cuobjdump -sass of covasim_b200/libcovasim_b200.so, sm_100a; written by profiles/sass_excerpts.py).',
       'What to look for: LDG.E.128 = 128-bit global loads (edge quads: p1 / p2 / beta of four edges per load; adjacency entries and agent records: 16 bytes',
       'each); .CONSTANT = the non-coherent path (only data that is never written during a run: the adjacency, the edge lists of the dense pass);',
       '.STRONG.GPU = L2 loads (ld.global.cg) of data the previous kernel of the chain wrote (the kernel may have been resident already: programmatic',
       'dependent launch); CCTL / PREFETCH = the L2 prefetch of the dense pass; ACQBULK = the programmatic-dependent-launch wait (griddepcontrol.wait).',
       'UBLKCP (cp.async.bulk global -> shared) and SYNCS (mbarrier) appear only in edge_pass_tma_kernel, the bulk-copy-staged form of the dense pass that was',
       'built, measured slower and kept as a profiling variant; the default dense pass keeps its tiles in registers (DESIGN.md section 4).', '']
print('\n'.join(hdr + out))
