# GPU test suite + multi-form smoke against the -DRRS_DEBUG_CHECKS build (device-side asserts on every stack push,
# queue slot, node / primitive / pixel index).  Build it first:
#   RRS_NVCC_EXTRA="-DRRS_DEBUG_CHECKS" python -m rayrs_b200.build --force && mkdir -p _variants && cp rayrs_b200/librayrs_b200.so _variants/lib_debugchecks.so && python -m rayrs_b200.build --force
cp rayrs_b200/librayrs_b200.so /tmp/orig.so
cp _variants/lib_debugchecks.so rayrs_b200/librayrs_b200.so
python scripts/sanitize_small.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size_config and not multigpu and not config5 and not bvh_build" 2>&1 | tail -3
timeout 200 python scripts/gpu_dev.py c4,c5 0 4 2>&1 | grep -v "scene build"
cp /tmp/orig.so rayrs_b200/librayrs_b200.so
