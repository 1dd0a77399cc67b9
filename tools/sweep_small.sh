for n in 2368 8704 10001 18944; do
  for K in 1 2; do
    for shape in 1 2 4; do
      B200BLS_KERNEL=$K python tools/prof_run.py pairing $n $shape 2
    done
  done
done
