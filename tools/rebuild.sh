#!/bin/bash
# incremental build of krotov_b200/csrc/libkrotov_b200.so (same flags as __graft_entry__.build)
cd "$(dirname "$0")/.." && python -c "
from krotov_b200 import _lib
_lib.build_library(verbose=False)
print('built', _lib.LIB_PATH if hasattr(_lib, 'LIB_PATH') else '')
" 2>&1 | tail -3
