# isolated synchronous batches (automatic shape) under each forced shape of the one-thread kernel, and the new default
for n in 37888 56832 65536 75776 100000 131072 151552; do
  for S in 2 4 5; do
    echo -n "forced shape $S: "; B200BLS_KERNEL=1 B200BLS_ISOLATED_SHAPE=$S python tools/prof_run.py pairing $n 0 2
  done
  echo -n "default policy: "; python tools/prof_run.py pairing $n 0 2
done
# whole waves, explicit shapes (CTA-wide item blocks)
python tools/prof_run.py pairing 56832 4 2
python tools/prof_run.py pairing 75776 5 2
python tools/prof_run.py verify 56832 4 2
python tools/prof_run.py verify 75776 5 2
