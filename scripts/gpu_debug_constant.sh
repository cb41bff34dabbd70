cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "textureless" -s 2>&1 | grep -E "worst|FAILED|passed|failed|max err" | head -12; done
python - <<'PY'
import sys; sys.path.insert(0,'tests')
import numpy as np
import localexpstereo_b200 as L
from oracle import lexp_oracle as O
import test_gpu_parity as T
H, W, D, windR = 120, 150, 16, 20
g = T._textureless_guides(H, W)["constant"]
volL = O.synthetic_volume(D, H, W, 77)
prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
E = L.CostVolumeEnergy(g, None, volL, None, prm, D - 1)
Or = O.CostVolumeEnergyOracle(g, None, volL, None, windR, 1e-4, 0.5, D - 1)
rng = O.CvRNG(5)
for (f, t) in [((0, 0, 100, 90), (0, 0, 60, 50)), ((30, 20, 120, 100), (50, 40, 80, 60)), ((60, 40, 90, 80), (80, 60, 50, 40))]:
    for _ in range(3):
        p = O.create_random_label(rng, t[0] + 5, t[1] + 5, 0, D - 1)
        outs=[]
        for rep in range(3):
            img = np.zeros((H, W), np.float32)
            E.ComputeUnaryPotential(f, t, img[f[1]:f[1] + f[3], f[0]:f[0] + f[2]], p)
            outs.append(img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]].copy())
        ref = Or.compute_unary_potential(f, t, p)
        ok = ref != O.COST_FOR_INVALID
        err = np.abs(outs[0][ok]-ref[ok])/np.maximum(np.abs(ref[ok]),1e-3)
        print(f, t, 'max rel err %.3e'%err.max(), 'repeatable', np.array_equal(outs[0],outs[1]) and np.array_equal(outs[0],outs[2]), 'n bad', int((err>1e-4).sum()), 'of', int(ok.sum()))
        if err.max()>1e-4:
            bad=np.argwhere((np.abs(outs[0]-ref)/np.maximum(np.abs(ref),1e-3)>1e-4)&ok)
            print('  bad rows', np.unique(bad[:,0])[:20], 'cols', np.unique(bad[:,1])[:20], 'got', outs[0][bad[0][0],bad[0][1]], 'ref', ref[bad[0][0],bad[0][1]])
E.close()
PY
