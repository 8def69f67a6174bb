#!/bin/sh
# Builds the host-emulation DEBUGGING harness (not shipped; see host_emul.cu header).
set -e
cd "$(dirname "$0")"
nvcc -O2 -std=c++17 -arch=sm_100a -x cu -Xcompiler -fPIC -shared -o libk2b_host_emul.so host_emul.cu
nvcc -O2 -std=c++17 -arch=sm_100a -x cu -Xcompiler -fPIC -shared -o libk2b_warp_emul.so warp_emul.cu
nvcc -O2 -std=c++17 -arch=sm_100a -x cu -Xcompiler -fPIC -shared -o libk2b_artic_emul.so artic_emul.cu
