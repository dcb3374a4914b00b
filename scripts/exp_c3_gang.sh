# pass B without the cache on C3-shaped input (the per-GPU share of bench.py's C3): CTA size
C3="--reads 12500000 --genome 12500000 --ksize 21 --sub-ppm 10000 --n-ppm 1000 --steps 2 --hint 280000000"
for th in 768 512 1024; do
  echo "== OXLI_B200_AGG_THREADS=$th"
  OXLI_B200_AGG_THREADS=$th timeout 100 python scripts/exp_part.py $C3 --digest --configs fused,auto
done
echo "== C2 (regression check)"
timeout 100 python scripts/exp_part.py --digest --configs fused,auto
