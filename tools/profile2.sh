#!/bin/bash
# usage (on the GPU box): tools/profile2.sh LABEL KEY REGEX SKIP GAMES -- <bench.py arguments>
# 1. the bench command without ncu (must exit 0; its JSON line gives the game-cycles per launch)  2. the launch list
# 3. one `--set full` capture of the (SKIP+1)-th launch matching REGEX  4. summaries: gpurun_out/LABEL_KEY_ncu_full.txt and an
#    entry KEY in gpurun_out/traffic_LABEL.json (tools/ncu_summary.py)
L=$1; K=$2; RX=$3; SKIP=$4; GAMES=$5; shift 5; [ "$1" = "--" ] && shift
CMD="python bench.py --no-cpu-baseline --no-e2e --no-secondary --prewarm-seconds 0 $*"
O=gpurun_out/${L}_${K}
$CMD > ${O}_plain.log 2> ${O}_plain.err || { echo "plain run failed ($K)"; tail -5 ${O}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv $CMD > ${O}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c 1 -f -o ${O}_prof $CMD > ${O}_ncu2.log 2>&1
ncu -i ${O}_prof.ncu-rep --page raw --csv > ${O}_raw.csv 2>/dev/null
ncu -i ${O}_prof.ncu-rep --page source --print-source sass --csv > ${O}_sass.csv 2>/dev/null
CYC=$(python -c "import json,sys; d=json.loads(open('${O}_plain.log').read().strip().splitlines()[-1]); print(d['stats']['window_game_cycles']/d['steps'])")
python profiles/ncu_raw_to_txt.py ${O}_raw.csv "ncu --set full --clock-control none -k regex:$RX -s $SKIP -c 1 of \`$CMD\` ($CYC game-cycles per launch)" > ${O}_ncu_full.txt
python tools/ncu_summary.py ${O}_raw.csv $K "ncu --set full capture profiles/${L}_${K}_ncu_full.txt: \`bench.py $*\`, launch $((SKIP+1)) matching $RX" $CYC $GAMES gpurun_out/traffic_$L.json
rm -f ${O}_prof.ncu-rep
