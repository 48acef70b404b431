#!/bin/bash
# Same-box A/B: tools/ab_run.sh <variant> <variant> ...  where a variant is "main", a library tag built by
# tools/ab_build.sh, or "env:NAME=VALUE[,NAME=VALUE]" (kernel-selection switches of the in-tree library).  Alternates the
# variants twice; prints per-level ResBlock milliseconds (best of 3 forwards at bench size) and the sustained forward time.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for round in 1 2; do
  for TAG in "$@"; do
    ENVS=""
    unset VT_LIB_PATH
    case "$TAG" in
      main) ;;
      env:*) ENVS=$(echo "${TAG#env:}" | tr ',' ' ') ;;
      *) export VT_LIB_PATH=$PWD/build/ab/libvocalie_b200_$TAG.so ;;
    esac
    env $ENVS timeout 200 python tools/hift_timeline.py --reps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())['ms']
l = [d['source_resblock%d' % i] + d['resblocks%d' % i] for i in range(3)]
print('$TAG round $round: L0 %.3f L1 %.3f L2 %.3f (src %.3f rb %.3f) rb %.3f total %.3f' % (l[0], l[1], l[2], d['source_resblock2'], d['resblocks2'], sum(l), d['total']))"
    env $ENVS timeout 200 python tools/hift_sustained.py 2>&1 | tail -1 | sed "s/^/$TAG /"
  done
done
