for mb in 2048 256 128 96 64 48; do echo "WS_MB=$mb"; DYNPY_FFT_WS_MB=$mb python tools/bench_other.py 4096 2>&1 | grep "^QR"; done
