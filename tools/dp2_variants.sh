for v in t512_k2 t512_k4 t1024_k2 t1024_k1; do
  DAB_LIB_PATH=$PWD/gpurun_variants/lib_$v.so timeout 150 python tools/profile_pair.py --reps 3 > gpurun_out/r2s27_pair_$v.json 2>/dev/null
  echo $v; tail -c 1400 gpurun_out/r2s27_pair_$v.json | grep -o '"dp2_trace": [0-9.]*\|"dp2": [0-9.]*\|"n_path2": [0-9]*\|"n_dp2_queries": [0-9]*\|"features_video": [0-9.]*\|"features_audio": [0-9.]*' | tr '\n' ' '; echo
done
timeout 150 python tools/profile_pair.py --reps 3 > gpurun_out/r2s27_pair_base.json 2>/dev/null
echo base; tail -c 1400 gpurun_out/r2s27_pair_base.json | grep -o '"dp2_trace": [0-9.]*\|"dp2": [0-9.]*\|"n_path2": [0-9]*\|"n_dp2_queries": [0-9]*\|"features_video": [0-9.]*\|"features_audio": [0-9.]*' | tr '\n' ' '; echo
