#!/bin/bash
# A/B of two builds of the library on the bench's headline numbers: tools/gpu_libswap.sh tools/bin/other.so
cp torchoptics_b200/libtorchoptics_b200.so /tmp/lib_a.so
for rep in 1 2; do
  cp /tmp/lib_a.so torchoptics_b200/libtorchoptics_b200.so; echo -n "A (in-tree)  "; bash tools/gpu_kernel_ms.sh
  cp $1 torchoptics_b200/libtorchoptics_b200.so; echo -n "B ($1)  "; bash tools/gpu_kernel_ms.sh
done
cp /tmp/lib_a.so torchoptics_b200/libtorchoptics_b200.so
