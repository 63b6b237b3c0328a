import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import time
import numpy as np
from emu_lib import emu_compressor
from oracle_lib import oracle_compress
from smallz4_b200 import corpus
BS=131072
c = emu_compressor(block_size=BS, batch_blocks=2)
bad=0
def chk(tag,data,lvl):
    global bad
    out = c.compress(data, level=lvl)
    ref,st = oracle_compress(data, lvl, block_size=BS)
    ok = out==ref; bad += (not ok)
    print(tag,lvl,len(data),len(out),len(ref),'OK' if ok else 'MISMATCH',f'redos={c.last_dp_redos()},{c.last_path_redos()}',flush=True)
rnd=lambda n,s: corpus.make('random',n,s).tobytes()
txt=lambda n,s: corpus.make('text',n,s).tobytes()
binr=lambda n,s: corpus.make('binary',n,s).tobytes()
for lvl in [9,5,1,3]:
    chk('zeros+random', bytes(50000)+rnd(30000,1)+bytes(40000)+rnd(9000,2)+txt(30000,3), lvl)
    chk('text+random', txt(40000,1)+rnd(20000,1)+txt(40000,2)+rnd(20000,3)+txt(30000,4), lvl)
    chk('binary+random', binr(40000,1)+rnd(20000,1)+binr(40000,2)+rnd(5000,3)+binr(30000,4), lvl)
    chk('runs+random', corpus.make('runs',60000,3).tobytes()+rnd(20000,1)+corpus.make('runs',60000,4).tobytes()+rnd(7000,5)+b'\x09'*70000+rnd(6000,6)+txt(8000,1), lvl)
    for seed in [1,2,3,4]:
        chk('mixed',corpus.make('mixed',400000,seed).tobytes(),lvl)
chk('random',rnd(200000,9),9)
chk('random-tail',txt(70000,9)+rnd(250000,9),9)
chk('random-gaps',rnd(100000,1)+txt(300,2)+rnd(70000,3)+txt(3000,2)+rnd(90000,4),9)
for lvl in ([] if 'quick' in sys.argv else [9]):   # long runs are slow in the emulator (minutes each)
    chk('zeros+text', bytes(120000)+txt(30000,5)+b'\x07'*100000+txt(12144,6), lvl)
    chk('zeros-rand-zeros', bytes(90000)+rnd(100,3)+bytes(90000)+rnd(50,4)+bytes(81994), lvl)

print('BAD',bad)
