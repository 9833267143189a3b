for c in 9 10 11 12 13; do MSM_G2=1 MSM_TABLE=$c python tools/tune_msm.py 13 $c $c | cut -c1-200; done
echo plain; MSM_G2=1 python tools/tune_msm.py 13 7 11 | cut -c1-200
echo g1; for c in 10 12 13; do MSM_TABLE=$c python tools/tune_msm.py 13 $c $c | cut -c1-200; done
