#!/bin/bash
# SASS opcode census per kernel family of the in-tree library (cuobjdump, no GPU needed):
# which Blackwell instructions the kernels really contain.
cd "$(dirname "$0")/.."
LIB=krotov_b200/csrc/libkrotov_b200.so
OUT=${1:-profiles/r02/sass_census.txt}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
python3 - "$TMP" "$OUT" <<'P'
import re, sys, collections
src, out = sys.argv[1], sys.argv[2]
fam = collections.defaultdict(collections.Counter)
kern = None
names = {}
for line in open(src, errors='replace'):
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        k = m.group(1)
        f = ('k_krotov_picard' if 'k_krotov_picard' in k else
             'k_dp_sweep' if 'k_dp_sweep' in k else
             'k_dp_build' if 'k_dp_build' in k else
             'k_dp_(plan|segprod|expand|epilogue)' if 'k_dp_' in k else
             'k_fwupd_sat' if 'k_fwupd_sat' in k else
             'k_fwupd_rows / k_prop_rows / k_rows_prep' if '_rows' in k and 'k_seg_chain' not in k else
             'k_sweep_csr' if 'k_sweep_csr' in k else
             'k_sweep_warp' if 'k_sweep_warp' in k else
             'k_fwupd_spec' if 'k_fwupd_spec' in k else
             'k_prop_spec / k_seg_chain' if ('k_prop_spec' in k or 'k_seg_chain' in k) else
             'k_prop_small / k_fwupd_small' if '_small' in k else 'other')
        kern = f
        names.setdefault(f, set()).add(k)
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1)
        fam[kern][op.split('.')[0]] += 1
        if op.startswith(('UBLKCP', 'SYNCS', 'UTMA', 'LDGSTS', 'REDUX', 'ACQBULK', 'PREEXIT',
                          'DMMA', 'UTC', 'LDTM', 'SHFL', 'BAR', 'WARPSYNC', 'LDGDEPBAR')):
            fam[kern][op] += 0
interesting = ['DFMA', 'DADD', 'DMUL', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'WARPSYNC',
               'UBLKCP', 'SYNCS', 'LDGSTS', 'CREDUX', 'ACQBULK', 'PREEXIT', 'LDL', 'STL',
               'UTMALDG', 'DMMA', 'HMMA', 'UTCHMMA', 'LDTM', 'MUFU', 'REDG', 'ATOMG']
with open(out, 'w') as fh:
    fh.write("SASS census of krotov_b200/csrc/libkrotov_b200.so (cuobjdump -sass, sm_100a), static instruction\n"
             "counts summed over all instantiations of a kernel family.  UBLKCP = cp.async.bulk (TMA bulk copy),\n"
             "SYNCS = mbarrier ops, LDGSTS = cp.async, CREDUX = warp integer reduction (redux.sync), ACQBULK / PREEXIT =\n"
             "griddepcontrol.wait / launch_dependents, LDL / STL = local-memory (spill) traffic, REDG = reductions performed at L2\n"
             "(red.global.max.u64: slot publication of k_fwupd_sat).  There is no\n"
             "tensor-core instruction (DMMA / UTC*MMA / LDTM): the arithmetic is complex128 on the FP64 pipe.\n\n")
    fh.write("%-38s %6s " % ("family", "#inst") + " ".join("%8s" % c for c in interesting) + "\n")
    for f in sorted(fam):
        c = fam[f]
        fh.write("%-38s %6d " % (f, len(names[f])) + " ".join("%8d" % c.get(k, 0) for k in interesting) + "\n")
print(open(out).read())
P
rm -f "$TMP"
