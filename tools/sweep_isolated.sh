# isolated synchronous batches (automatic shape) under each forced shape of the one-thread kernel, and the default policy
for n in 56832 60000 65536 75776 100000 131072 151552; do
  for S in 4 5; do
    echo -n "forced shape $S: "; B200BLS_KERNEL=1 B200BLS_ISOLATED_SHAPE=$S python tools/prof_run.py pairing $n 0 2
  done
  echo -n "default policy: "; python tools/prof_run.py pairing $n 0 2
done
