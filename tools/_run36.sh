cd $GRAFT_REPO_ROOT
for c in pair pair-ragged old old-ragged ws bwd; do timeout 600 compute-sanitizer --tool racecheck python tools/race_probe.py $c 2>&1 | grep "^ok\|RACECHECK SUMMARY\|Race reported" | sort | uniq -c | head -5; done
