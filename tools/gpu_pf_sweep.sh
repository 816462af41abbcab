for pa in 1 0 2; do for px in 1 0 2 3; do
  echo -n "PF_A=$pa PF_X=$px: "; DSEN2_PAIR_PF_A=$pa DSEN2_PAIR_PF_X=$px DSEN2_DIAG_N=84 timeout 100 python tools/gpu_diag.py --one pair_res32_alone 2>&1 | grep "RESIDUALQ :\|RELU" | awk '{printf "%s %s ms | ", $3, $5}'; echo
done; done
