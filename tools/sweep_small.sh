# isolated batches, automatic shape (ctas_per_sm = 0): one-thread kernel vs paired kernel
for n in 1 300 2368 8704 10001 18944 37888 56832 65536 131072; do
  for K in 1 2; do
    B200BLS_KERNEL=$K python tools/prof_run.py pairing $n 0 2
  done
done
B200BLS_KERNEL=2 python tools/prof_run.py pairing 56832 4 2
B200BLS_KERNEL=1 python tools/prof_run.py pairing 56832 4 2
