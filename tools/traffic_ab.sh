ARGS="--columns 151552 --hours 24 --steps 2 --warmup 1 --no-cpu-baseline --no-steady --no-extra --spin 20"
for N in "$@"; do
  MC_B200_LIB=$(realpath build/ab/$N.so) ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_transient_fast -s 1 -c 1 --csv python bench.py $ARGS 2>/dev/null | grep -E "dram__|gpu__time|lts__" | awk -F'","' -v n=$N '{print n, $(NF-2), $(NF-1), $NF}'
done
