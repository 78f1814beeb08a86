#!/bin/bash
# Build a variant of libeegfe.so for A/B timing with tools/kbench.py (development tool).
#   tools/build_variant.sh <git-rev|WORK> <out.so> [extra nvcc flags...]
# <git-rev>: take eeg2video_b200/csrc/* from that revision; WORK: the working tree.
set -e
rev=$1; out=$2; shift 2
mkdir -p "$(dirname "$out")"
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d "$root/eeg2video_b200/csrc/_var.XXXXXX")
trap 'rm -rf "$tmp"' EXIT
if [ "$rev" = WORK ]; then
  cp "$root"/eeg2video_b200/csrc/*.cu "$root"/eeg2video_b200/csrc/*.cuh "$root"/eeg2video_b200/csrc/*.h "$tmp"/
else
  for f in $(git -C "$root" ls-tree --name-only "$rev" eeg2video_b200/csrc/ | xargs -n1 basename | grep -E "\.(cu|cuh|h)$"); do git -C "$root" show "$rev:eeg2video_b200/csrc/$f" > "$tmp/$f"; done
fi
mkdir -p "$tmp/inc"; 
if [ "$rev" = WORK ]; then cp "$root/include/eegfe.h" "$tmp/inc/"; else git -C "$root" show "$rev:include/eegfe.h" > "$tmp/inc/eegfe.h"; fi
sed -i 's#"../../include/eegfe.h"#"inc/eegfe.h"#' "$tmp/eegfe_kernels.cu"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC -shared "$@" -o "$out" "$tmp/eegfe_kernels.cu"
echo "built $out"
