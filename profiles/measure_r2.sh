set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.err
python profiles/kernel_bench.py > gpurun_out/r2_kernels.json 2> gpurun_out/r2_kernels.txt
python profiles/config_bench.py > gpurun_out/r2_configs64.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"decode_filter_pairs|nms_select|nms_kernel|lb_general|detect_decode|filter_pred_dense" -s 12 -c 6 -o gpurun_out/r2x_prof python profiles/run_one.py 64 > gpurun_out/r2x_prof.log 2>&1
tail -n 3 gpurun_out/r2x_prof.log
