import sys
sys.path.insert(0,'/root/repo')
from smallz4_b200 import corpus
from smallz4_b200.api import Compressor
c=Compressor(device=0)
d=corpus.make(sys.argv[1], 64<<20, 1)
c.compress(d, level=9)
