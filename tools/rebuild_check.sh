#!/bin/bash
# development helper: rebuild the engine and report spills + whether the scan hot loop kept its uniform-register operands
cd "$(dirname "$0")/.." || exit 1
python -m spotify_recommender_b200.build -v 2>&1 | grep -E "error|spill" | sort | uniq -c
for v in ILi8ELi256ELi2ELb1ELi1ELb1E ILi8ELi256ELi2ELb1ELi2ELb1E ILi8ELi256ELi2ELb1ELi0ELb1E ILi8ELi512ELi1ELb1ELi1ELb1E ILi8ELi512ELi1ELb1ELi0ELb0E ILi8ELi512ELi1ELb1ELi0ELb1E; do
  cuobjdump -sass spotify_recommender_b200/libsr_engine.so | awk "/Function : .*scan_kernel$v/{p=1} p&&/Function : /&&!/scan_kernel$v/{p=0} p" > /tmp/scan_$v.sass
  echo $v lines=$(grep -c "" /tmp/scan_$v.sass) ffma2=$(grep -c FFMA2 /tmp/scan_$v.sass) ffma2_ur=$(grep FFMA2 /tmp/scan_$v.sass | grep -c "UR") calls=$(grep -c CALL /tmp/scan_$v.sass)
done
