set -x
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 > gpurun_out/b.json 2> gpurun_out/b.err
ncu --set full --clock-control none --import-source on -k regex:SmemPacked8 -s 51 -c 2 -o gpurun_out/prof_bsw_r1f_left python bench.py --steps 1 --warmup 2 > gpurun_out/ncu_full_bsw1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:SmemPacked8 -s 59 -c 2 -o gpurun_out/prof_bsw_r1f_right python bench.py --steps 1 --warmup 2 > gpurun_out/ncu_full_bsw2.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:smem_bwd_kernel|smem_fwd1_kernel|tail_kernel" -s 9 -c 4 -o gpurun_out/prof_smem_r1f python bench.py --steps 1 --warmup 2 > gpurun_out/ncu_full_smem.log 2>&1
ls -la gpurun_out
