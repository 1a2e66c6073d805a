bash tools/_typed2.sh
for path in strip wide; do for inst in 10000 1000 128 64; do
  echo "== $path $inst" >> gpurun_out/r02_rand.log
  P3CS_PATH=$path timeout 200 python tools/kbench.py --configs c3rand --instances $inst >> gpurun_out/r02_rand.log 2>&1
done; done
sed 's/"vertices".*//' gpurun_out/r02_rand.log
