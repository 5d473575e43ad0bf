#!/bin/bash
# build_variant.sh <commit> <name>: build that commit's csrc into build/variants/lib<name>.so (for same-box A/B runs)
set -e
mkdir -p build/variants
rm -rf /tmp/zest_wt_$2 && git worktree add -f /tmp/zest_wt_$2 $1 > /dev/null 2>&1
(cd /tmp/zest_wt_$2 && python __graft_entry__.py > /dev/null && cp zest_nerf_b200/libzest_b200.so /root/repo/build/variants/lib$2.so)
git worktree remove --force /tmp/zest_wt_$2
ls -la build/variants/lib$2.so
