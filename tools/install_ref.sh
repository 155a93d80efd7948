#!/bin/bash
# Copies the five reference files of the render path, UNMODIFIED, into the git-ignored oracle/_ref/ so that the reference
# itself (not a port) can be timed as the CPU arm on the GPU box, where /root/reference does not exist.
# oracle/_ref/ is listed in .gitignore (reference sources never enter the history) and not in .gpurunignore (it travels).
# Run by __graft_entry__.build() whenever /root/reference is present.
set -e
REF=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/oracle/_ref"
[ -d "$REF/nerf" ] || { echo "install_ref: $REF not found (nothing to do)"; exit 0; }
mkdir -p "$DST/nerf" "$DST/pi_GAN"
cp "$REF/nerf/render.py" "$REF/nerf/nerf.py" "$DST/nerf/"
cp "$REF/pi_GAN/render.py" "$REF/pi_GAN/modules.py" "$REF/pi_GAN/utils.py" "$DST/pi_GAN/"
( cd "$REF" && sha256sum nerf/render.py nerf/nerf.py pi_GAN/render.py pi_GAN/modules.py pi_GAN/utils.py ) > "$DST/SHA256SUMS"
echo "install_ref: reference render path copied to $DST"
