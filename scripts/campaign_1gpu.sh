# one-GPU measurement campaign of the round: every config's bench line + the ncu captures profiles/ keeps
set -x
O=gpurun_out/camp
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/c3_quicked.json 2> $O/c3_quicked.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/c3_ref.json 2> $O/c3_ref.err
python bench.py --workload c3 --algo banded --steps 5 --warmup 3 > $O/c3_banded.json 2> $O/c3_banded.err
python bench.py --workload c3 --algo windowed --steps 5 --warmup 3 > $O/c3_windowed.json 2> $O/c3_windowed.err
python bench.py --workload c2 --steps 10 --warmup 3 > $O/c2_quicked.json 2> $O/c2_quicked.err
python bench.py --workload c1 --steps 20 --warmup 3 > $O/c1_quicked.json 2> $O/c1_quicked.err
python bench.py --workload c4 --steps 3 --warmup 3 > $O/c4_quicked.json 2> $O/c4_quicked.err
python bench.py --workload c4 --algo hirschberg --steps 3 --warmup 3 > $O/c4_hirschberg.json 2> $O/c4_hirschberg.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/c5_quicked.json 2> $O/c5_quicked.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_quicked.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-packed > $O/launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 22 --launch-count 22 -o $O/ncu_c3_quicked -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-packed > $O/ncu.log 2>&1
ncu -i $O/ncu_c3_quicked.ncu-rep --page raw --csv > $O/ncu_c3_quicked_raw.csv 2>/dev/null
rm -f $O/ncu_c3_quicked.ncu-rep
tail -c 200 $O/*.err
