for sg in 32 16 8; do echo "sg=$sg"; TES_TC5_SG=$sg python scripts/dev_tc5_time.py 1000000 6 1024 1024 1 1 0:-4 2>&1 | tail -2; done
