for t in 0; do echo "== KFSP_BOX_TUNE=$t"; KFSP_BOX_TUNE=$t ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_spmv_bd2 --csv python tools/lattice_tune.py one 1 10000 10000 2>/dev/null | grep -E "k_spmv_bd2" | python -c "
import sys,csv,collections
agg=collections.defaultdict(list)
for r in csv.reader(sys.stdin):
    if len(r)<15: continue
    agg[r[4][5:34]].append(float(r[14].replace(',','')))
for k,v in sorted(agg.items()): print('  ',k, len(v), round(sum(v)/len(v)/1e3,1),'us')
"; done
