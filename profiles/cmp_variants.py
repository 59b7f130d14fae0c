import numpy as np, sys, glob
base=np.load('gpurun_out/ret_w1.npy')
for f in sorted(glob.glob('gpurun_out/ret_w*.npy')):
    r=np.load(f); print(f, 'bitwise' if np.array_equal(r,base) else 'max rel %.3g'%np.max(np.abs(r-base)/np.abs(base)))
