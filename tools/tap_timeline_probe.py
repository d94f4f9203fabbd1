# NOTE: the in-kernel globaltimer stamps are compiled in only with -DSTGCN_TIMELINE:
#   tools/build_variants.sh tl="-DSTGCN_TIMELINE" && STGCN_B200_LIB=$PWD/build/variants/tl.so python <this script>
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, stgcn_b200
from stgcn_b200 import layers, _lib as L
stgcn_b200.set_precision("bf16")
dev=torch.device('cuda')
buf=torch.zeros(32,dtype=torch.int64,device=dev)
def run(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    L.check(L.lib().stgcn_debug_timeline(buf.data_ptr()))
    buf.zero_(); fn(); torch.cuda.synchronize()
    L.check(L.lib().stgcn_debug_timeline(None))
    t=buf.cpu().tolist(); t0=t[0]
    lab=['start','setup done','weights ready(mma)','tile0 full(mma)','tile0 tfull(epi)','tile0 stored','epi done','end','tile8 full(mma)','tile8 tfull(epi)','tile16 tfull(epi)']
    print(name, {lab[i]: (t[i]-t0)/1e3 if t[i] else None for i in range(11)})
    lab3={16:'mma t8 begin',17:'tempty ok',18:'full j0',19:'full j1',20:'full j2',22:'mmas issued',23:'tfull committed',24:'stages released'}
    print('   MMA thread, tile 8 (us):', {v: (t[k]-t0)/1e3 if t[k] else None for k,v in lab3.items()})
    lab4={25:'prod g24 begin',26:'empty ok',27:'issued',28:'g24 done / g25 issued',29:'g32 issued'}
    print('   producer (us):', {v: (t[k]-t0)/1e3 if t[k] else None for k,v in lab4.items()})
    lab2={9:'tile8 tfull',11:'t8 chunk0 start',12:'t8 chunk0 tmem loaded',13:'t8 chunk0 math done',14:'t8 all chunks done',15:'t8 released+store issued'}
    print('   tile 8 detail (us):', {v: (t[k]-t0)/1e3 if t[k] else None for k,v in lab2.items()})
B,N=256,228
for (cin,cout,T) in [(16,64,10),(64,64,8)]:
    lay=layers.TemporalConvLayer(3,cin,cout,N,'glu').to(dev)
    x=torch.randn(B,cin,T,N,device=dev)
    run(f"tconv {cin}->{cout} T={T}", lambda: lay(x))
gl=layers.GraphConvLayer('cheb_graph_conv',64,16,3,torch.eye(N,device=dev),True).to(dev)
x=torch.randn(B,64,10,N,device=dev)
run("gconv (last tap launch = mix)", lambda: gl(x))
