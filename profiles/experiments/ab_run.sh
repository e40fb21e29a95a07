timeout 900 python profiles/experiments/r2_sweep.py 1:0,f1x1,f1x7 256:0,f1x7,f1x1 4096:0,f14x7 8192:0,f28x7,f28x5,f32x7 16384:0,f28x7,f32x7 65536:0,f32x7
