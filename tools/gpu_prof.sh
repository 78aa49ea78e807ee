#!/bin/bash
# ncu --set full captures of the two top decoder kernels (third pass of tools/prof_decoder.py = warm).
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused_kernel -s 4 -c 1 \
    -o $OUT/prof_fused python tools/prof_decoder.py 3 > $OUT/ncu_fused.log 2>&1; echo "fused exit $?"
# pass 3 linear launches: skip 2*N_LINEAR; capture lin_z, fc_0, fc_1 of block 0 and the first Qa (832 wide)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s ${SKIP_LIN:-49} -c 3 \
    -o $OUT/prof_linear python tools/prof_decoder.py 3 > $OUT/ncu_linear.log 2>&1; echo "linear exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_decoder.csv \
    python tools/prof_decoder.py 3 > $OUT/ncu_launches.log 2>&1; echo "launch list exit $?"
ls -la $OUT
