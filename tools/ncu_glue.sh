#!/bin/bash
# ncu --set full captures of the non-convolution kernels of one step (B = 64, W48) and of the LBS route at B = 8192
# usage (on the GPU box): bash tools/ncu_glue.sh
set -x
for k in k_smpl_verts k_smpl_pose k_smpl_joints k_faces k_resolve k_project k_stn_sample k_stn_params k_gcn_pose_head k_fuse_sum k_iuv_clean_global k_iuv_clean_parts k_maxpool3x3s2; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_ncu_$k python bench.py --steps 1 --warmup 1 --no-cpu --no-extras --no-graph > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none -k regex:k_smpl_verts -s 6 -c 1 -f -o gpurun_out/r02_ncu_k_smpl_verts_b8192 python tools/lbs_sweep.py one > gpurun_out/ncu_lbs8192.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_conv_tc -s 6 -c 1 -f -o gpurun_out/r02_ncu_lbs_gemm_b8192 python tools/lbs_sweep.py one >> gpurun_out/ncu_lbs8192.log 2>&1
