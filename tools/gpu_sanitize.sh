#!/bin/bash
# compute-sanitizer memcheck over one small invocation of every kernel (smoke) -- one tool per gpurun call
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 &&
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_memcheck.log 2>&1
echo "exit $?"; tail -5 gpurun_out/sanitizer_memcheck.log; tail -2 gpurun_out/smoke_memcheck.log
