#!/bin/bash
# SURVEY 8f.2 probe: is hardware video decode (NVDEC) reachable from this image on the GPU box?
echo "== ldconfig"; ldconfig -p | grep -i -E "nvcuvid|nvidia-encode|nvjpeg|avcodec" 
echo "== files"; ls -la /usr/lib/x86_64-linux-gnu/libnvcuvid* /usr/lib64/libnvcuvid* 2>&1 | head
echo "== headers"; find / -name "nvcuvid.h" -o -name "cuviddec.h" 2>/dev/null | head
echo "== python"; python - <<'PY'
import importlib
for m in ("PyNvVideoCodec", "torchcodec", "decord", "av", "nvidia.dali", "torchvision.io"):
    try:
        importlib.import_module(m); print(m, "importable")
    except Exception as e:
        print(m, "missing:", type(e).__name__)
try:
    import torchvision
    print("torchvision", torchvision.__version__, "has VideoReader:", hasattr(torchvision.io, "VideoReader"))
    from torchvision.io import _HAS_GPU_VIDEO_DECODER
    print("torchvision GPU video decoder built:", _HAS_GPU_VIDEO_DECODER)
except Exception as e:
    print("torchvision probe:", type(e).__name__, e)
import cv2
print("cv2 cudacodec:", hasattr(cv2, "cudacodec"))
PY
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
