#!/bin/bash
# Round 2, first GPU call: verified suite, the parked tests, headline bench + launch list, and full ncu captures of the
# diffusion variants of the step kernel (c4: RK4 + turb + meso; c3: RK4 + meso + sedi).
set -u
cd "$(dirname "$0")/.."
scripts/gpu_first_call.sh
WL=c4 NCU_SKIP=4 scripts/gpu_check.sh ncu
WL=c3 NCU_SKIP=4 scripts/gpu_check.sh ncu
ls -la gpurun_out/*.ncu-rep
