set -x
CMD="python bench.py --steps 2 --warmup 3 --no-analysis-leg --no-cpu-baseline --no-pcie-probe --no-verify"
$CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_fused -s 3 -c 1 -o gpurun_out/r2_k1 $CMD > gpurun_out/r2_ncu_k1.log 2>&1
TA_OVERLAP=0 python tools/all_kernels_once.py > gpurun_out/r2_allk_plain.log 2>&1 || exit 1
TA_OVERLAP=0 ncu --set full --clock-control none -k 'regex:chroma_project|onset_flux|time_domain_kernel|tempogram_sliding|pip_peaks|cqt_chroma|cqt_decimate|hpss_|true_peak|mfcc|nv_|tuning_kernel|ac_rows' -s 60 -c 40 -o gpurun_out/r2_allk python tools/all_kernels_once.py > gpurun_out/r2_ncu_allk.log 2>&1
ls -la gpurun_out/*.ncu-rep
