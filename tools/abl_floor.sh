for p in f16 1x 3x; do VSB_TC_CL=1 PREC=$p timeout 100 python tools/tc_ablation.py 1000000 1 5 9 0; done
