# round 2: compute-sanitizer memcheck over a small tour of every kernel variant (one tool, one small program) -- the tool answered "closed on this pool"; the plain run passed
O=gpurun_out; mkdir -p $O
timeout 300 python tests/kernel_tour.py > $O/san_plain.log 2>&1 && echo "plain ok" &&
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck.log python tests/kernel_tour.py > $O/san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 $O/san_plain.log; tail -5 $O/san_memcheck.log; tail -15 $O/memcheck.log
