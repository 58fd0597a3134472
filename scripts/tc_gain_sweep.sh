for g in 0,0 1,0 1,1 2,1 2,2 3,2 3,3; do
  echo "== gain $g"
  ADK_TC_GAIN=$g timeout 200 python -m pytest tests/test_gpu_parity.py -q -s -k "forward_matches and tc-simt and (jit2 or mixed or pbc_ttf or skew)" 2>&1 | grep "cuda-vs-fp64" | awk '{print $1, $3}' | tr '\n' ' '
  echo
done
