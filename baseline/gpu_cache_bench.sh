#!/bin/bash
# run on a GPU box from the repo root: reference gpu_cache vs hugectr_b200 cache kernels
./baseline/_ref/gpu_cache_bench hugectr_b200/lib/libhctr_cuda.so
