#!/bin/bash
# builds the product library and, beside it, the wait-profiler build (hcflow_b200/prof/, selected with HCFLOW_LIB=...)
set -e
cd "$(dirname "$0")/.."
HCF_BUILD_PROF=1 python -m hcflow_b200.build --force > /dev/null
mkdir -p hcflow_b200/prof
mv hcflow_b200/libhcflow_b200.so hcflow_b200/prof/libhcflow_b200_prof.so
python -m hcflow_b200.build --force > /dev/null
ls -la hcflow_b200/libhcflow_b200.so hcflow_b200/prof/libhcflow_b200_prof.so
