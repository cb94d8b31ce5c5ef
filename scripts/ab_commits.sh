#!/bin/bash
# Build libdpomp at older commits into lib/variants/libdpomp_<tag>.so (git worktrees under /tmp), for same-box A/B of the whole
# library across commits:  scripts/ab_commits.sh 24c4cc0:r2start c82c7df:comm ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/discretepomp.jl_b200/lib/variants"
for spec in "$@"; do
  c=${spec%%:*}; tag=${spec##*:}
  wt=/tmp/dpomp_wt_$tag
  rm -rf "$wt"; git -C "$ROOT" worktree prune; git -C "$ROOT" worktree add -f "$wt" "$c" > /dev/null 2>&1
  (cd "$wt" && python discretepomp.jl_b200/build.py > /dev/null)
  cp "$wt/discretepomp.jl_b200/lib/libdpomp.so" "$ROOT/discretepomp.jl_b200/lib/variants/libdpomp_$tag.so"
  git -C "$ROOT" worktree remove --force "$wt"
  echo "built $tag from $c"
done
