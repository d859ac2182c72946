for m in 0 1; do
CTCPS_PSI_SPLIT=$m timeout 170 ncu --set full --clock-control none --import-source on -k regex:k_psi_ --launch-skip 20 --launch-count 1 -o gpurun_out/r1w_C2_psi_s$m -f python bench.py --config C2 --steps 1 --warmup 1 --profile --single-mode --state lazy --no-cpu-baseline > gpurun_out/r1w_ncufull_C2_s$m.log 2>&1
echo "split=$m rc=$?"
done
ls -la gpurun_out/r1w*.ncu-rep
