timeout 600 python tools/sweep_configs.py --what cad120 --batches 32,64,128,256 2>&1 | grep cad120
