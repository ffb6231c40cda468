for lib in ""; do
  tag=$( [ -z "$lib" ] && echo new || echo head )
  if [ -n "$lib" ]; then export SMX_LIB=$PWD/$lib; else unset SMX_LIB; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ln_ --csv --log-file gpurun_out/r01q_ncu_ln_$tag.csv python tools/ncu_ln.py > /dev/null 2>&1
  echo "== $tag"; grep -E "gpu__time_duration" gpurun_out/r01q_ncu_ln_$tag.csv | awk -F'","' '{print $5, $(NF)}' | tail -8
done
