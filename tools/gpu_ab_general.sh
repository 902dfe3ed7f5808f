#!/bin/bash
# A/B of builds of the library on the asphere fused pass: tools/gpu_ab_general.sh tools/bin/a.so tools/bin/b.so ...
# (the in-tree library is measured first and last)
cp torchoptics_b200/libtorchoptics_b200.so /tmp/lib_tree.so
for rep in 1 2; do
  for lib in /tmp/lib_tree.so "$@"; do
    cp $lib torchoptics_b200/libtorchoptics_b200.so
    echo "== $lib"; python tools/profile_general.py 296 | cut -c1-90
  done
done
cp /tmp/lib_tree.so torchoptics_b200/libtorchoptics_b200.so
