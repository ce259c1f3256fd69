set -x
export QMHA_STREAM_QUANT=1
V=quantizedmha_b200/lib/variants
timeout 600 python tools/quant_ab.py stream_c4=quantizedmha_b200/lib/libqmha.so c3=$V/libqmha_sc3.so c3r128=$V/libqmha_sc3r128.so
unset QMHA_STREAM_QUANT
timeout 600 python tools/quant_ab.py cluster=quantizedmha_b200/lib/libqmha.so
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -x -k persistent 2>&1 | tail -2
