#!/bin/bash
# A/B of the cfg3 SGLD step against the tree of commit 02d3a7d (temporary worktree _old/, not committed)
for rep in 1 2; do
echo -n "old: "; (cd _old && python tools/bench_extra.py sgld 2>/dev/null | cut -c100-260)
echo -n "new: "; python tools/bench_extra.py sgld 2>/dev/null | cut -c100-260
done
echo -n "new FUSED_STEP=0: "; BHMC_FUSED_STEP=0 python tools/bench_extra.py sgld 2>/dev/null | cut -c100-260
