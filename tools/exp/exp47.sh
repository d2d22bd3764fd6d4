timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/memcheck_small.py > gpurun_out/r02f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r02f_memcheck.log
