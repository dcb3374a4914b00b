C3="--reads 12500000 --genome 12500000 --ksize 21 --sub-ppm 10000 --n-ppm 1000 --steps 2"
echo "== C3 shape per GPU, hinted"
timeout 150 python scripts/exp_part.py $C3 --hint 280000000 --digest --configs fused,auto,part:1024,part:2048
echo "== C3 shape per GPU, no hint, table kept between steps / fresh table every step"
timeout 150 python scripts/exp_part.py $C3 --hint 0 --digest --configs fused,auto
timeout 150 python scripts/exp_part.py $C3 --hint 0 --fresh --digest --configs fused,auto
echo "== C2 (regression check)"
timeout 100 python scripts/exp_part.py --digest --configs fused,auto,part:1024
