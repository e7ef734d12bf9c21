#!/bin/bash
# the measured spread of the double-build iteration counts against the oracle on the random domains
tag=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "random_domains and 8" > gpurun_out/${tag}_random_domains_double.txt 2>&1
grep "double-build" gpurun_out/${tag}_random_domains_double.txt | sort -k 9 -n | tail -12; tail -2 gpurun_out/${tag}_random_domains_double.txt
