# cycles per phase of k_step (variants/phases.so = a build with -DAGARCL_PHASE_TIMING); $1 = settle steps
cp variants/phases.so agarcl_b200/libagarcl_b200.so
timeout 600 python tools/exp_perstep.py ${1:-2000} 20 2>&1 | grep PHASES | tail -22 > gpurun_out/phases.txt
python - ${1:-2000} <<'PY'
rows=[list(map(int,l.split()[1:])) for l in open('gpurun_out/phases.txt')]
names=['kernel(per warp)','hash/virus cache+zero','pool: publish+batches','pool: barrier wait','player loop: lanes after the last tick_player','apply_removals(+b16)','barrier-4 wait','players_collision','move_foods/regen','prologue','epilogue','idle warps','tp: record+cell loads','tp: bot decision','tp: move+self collisions (not premoved)','tp: virus collisions','tp: pellets','tp: auto split+eat food','tp: emit/split/add','tp: recombine','tp: decay','tp: publish','player loop: rest before a tick_player','lanes: speculation (+ food re-check)','lanes: ordered commit']+['-']*6
d=[[b-a for a,b in zip(r0,r1)] for r0,r1 in zip(rows[:-1],rows[1:])]
n=len(d); tot=[sum(x[i] for x in d)/n for i in range(28)]
inst_total=sum(tot[1:11])
print('launches averaged',n,'  mean cycles per launch summed over warps')
for i in range(25): print(f'{names[i]:46s} {tot[i]/1e6:10.2f} Mcycles  {100*tot[i]/tot[0]:5.1f}% of warp-time')
# steps that contain a bot-decision tick (every 10th tick) against the others (the k-th printed line precedes launch k of the run)
import sys
settle=int(sys.argv[1]) if len(sys.argv)>1 else 2000
first=settle+20-len(d)-1 # exp_perstep: settle launches, then 20 timed ones; the last len(d) launches are in `d`
dec=[any((4*(first+k)+t)%10==0 for t in range(4)) for k in range(len(d))]
for flag,label in ((True,'steps WITH a decision tick'),(False,'steps without')):
    sel=[x for x,f in zip(d,dec) if f==flag]
    if not sel: continue
    t=[sum(x[i] for x in sel)/len(sel) for i in range(28)]
    print(label, len(sel), ' kernel Mcycles', round(t[0]/1e6,1))
    for i in range(1,25):
        if t[i]/t[0]>0.004: print(f'   {names[i]:46s} {t[i]/1e6:9.2f}')
PY
