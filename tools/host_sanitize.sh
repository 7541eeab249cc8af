#!/bin/bash
# Development aid: the host-side builders (spec JSON emitter, Pod metadata / ray start / template builders, kr_pod_build) compiled alone with
# AddressSanitizer + UndefinedBehaviorSanitizer and driven by their own test files.  No GPU.   usage: tools/host_sanitize.sh [pytest args]
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libkrhost_asan.so
( cd kuberay_b200/csrc && g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -shared -fPIC -o "$OUT" \
    kr_specjson.cpp kr_podmeta.cpp kr_raystart.cpp kr_raytemplate.cpp kr_podbuild.cpp )
ASAN_LIB=$(g++ -print-file-name=libasan.so)
LD_PRELOAD="$ASAN_LIB" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 KR_HOST_ONLY_LIB="$OUT" \
  python -m pytest tests/test_podmeta.py tests/test_raystart.py tests/test_raytemplate.py tests/test_podbuilder.py tests/test_spec_json.py -q -m "not gpu" -p no:cacheprovider "$@"
