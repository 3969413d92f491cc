#!/bin/bash
# One single-GPU box session: ncu --set full captures (with source) of the position kernel at config 3p, the quaternion
# path's closed-form call and the 65,536-configuration FK / Jacobian launches.  Each program runs plain first.
mkdir -p gpurun_out
python tools/prof_pos.py > gpurun_out/r2_prof_pos.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pos_retarget -s 1 -c 1 -f -o gpurun_out/r2_pos python tools/prof_pos.py >> gpurun_out/r2_prof_pos.log 2>&1
python tools/microbench.py --cases bq --bq-only "closed_form_dof+linkpos" --iters 3 > gpurun_out/r2_prof_bq.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:body_quat -s 3 -c 1 -f -o gpurun_out/r2_bq_noik python tools/microbench.py --cases bq --bq-only "closed_form_dof+linkpos" --iters 3 >> gpurun_out/r2_prof_bq.log 2>&1
python tools/prof_fk65536.py > gpurun_out/r2_prof_fk.log 2>&1 && \
HRT_PROF_SHORT=1 ncu --set full --clock-control none --import-source on -k regex:"jacobian|fk_limb" -s 1 -c 5 -f -o gpurun_out/r2_fk65536 python tools/prof_fk65536.py >> gpurun_out/r2_prof_fk.log 2>&1
tail -4 gpurun_out/r2_prof_pos.log gpurun_out/r2_prof_bq.log gpurun_out/r2_prof_fk.log
