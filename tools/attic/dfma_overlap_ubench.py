# does non-FP64 instruction issue overlap with the FP64 pipe?  64 DFMA (g_i = fma(c_i,a,g_i)) + N extra instrs per iteration
def kernel(name, n_int, n_lds, N=16):
    L=[];w=L.append
    w(f".visible .entry {name}(.param .u64 p_out, .param .u32 p_iters)")
    w(".maxntid 256, 1, 1")
    w("{")
    w(f".reg .f64 %c<{N}>, %g<{N}>, %a, %s, %ld<8>; .reg .u64 %po, %off; .reg .u32 %it, %n, %t, %r<3>, %i<16>, %sa; .reg .pred %p;")
    w(".shared .align 16 .b8 sm[4096];")
    w("ld.param.u64 %po, [p_out]; ld.param.u32 %n, [p_iters]; cvta.to.global.u64 %po, %po;")
    w("mov.u32 %r0, %ctaid.x; mov.u32 %r1, %ntid.x; mov.u32 %r2, %tid.x; mad.lo.u32 %t, %r0, %r1, %r2;")
    w("cvt.rn.f64.u32 %s, %t; mul.rn.f64 %s, %s, 0d3E10000000000000;")
    for i in range(N):
        w(f"add.rn.f64 %c{i}, %s, 0d{0x3FEFFFF000000000+i*4096:016X}; mov.f64 %g{i}, %s;")
    for i in range(16): w(f"mov.u32 %i{i}, %t;")
    w("add.rn.f64 %a, %s, 0d3FEFFFFF00000000; mov.u32 %sa, sm; st.shared.f64 [%sa], %s;")
    w("mov.u32 %it, 0;")
    w("LOOP:")
    ops=[]
    for rep in range(4):
        for i in range(N): ops.append(f"fma.rn.f64 %g{i}, %c{i}, %a, %g{i};")
    extra=[f"add.u32 %i{j%16}, %i{j%16}, %it;" for j in range(n_int)]+[f"ld.shared.v2.f64 {{%ld{(2*j)%8}, %ld{(2*j+1)%8}}}, [%sa+{16*(j%64)}];" for j in range(n_lds)]
    # interleave extras evenly
    step=max(1,len(ops)//max(1,len(extra))) if extra else 0
    out=[];ei=0
    for k,o in enumerate(ops):
        out.append(o)
        if extra and k%step==step-1 and ei<len(extra): out.append(extra[ei]); ei+=1
    out+=extra[ei:]
    for o in out: w(o)
    w("add.u32 %it, %it, 1; setp.lt.u32 %p, %it, %n; @%p bra LOOP;")
    w("mov.f64 %s, %a;")
    for i in range(N): w(f"add.rn.f64 %s, %s, %g{i};")
    for i in range(8): w(f"add.rn.f64 %s, %s, %ld{i};")
    for i in range(16): w(f"cvt.rn.f64.u32 %a, %i{i}; add.rn.f64 %s, %s, %a;")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off; st.global.f64 [%po], %s;")
    w("ret;"); w("}")
    return L
out=[".version 8.7",".target sm_100a",".address_size 64",""]
cfg=[('i0',0,0),('i16',16,0),('i32',32,0),('i64',64,0),('i128',128,0),('l8',0,8),('l16',0,16),('l32',0,32),('i32l16',32,16)]
for n,a,b in cfg: out+=kernel('ub_'+n,a,b)+[""]
open('build/ubench/ub2.ptx','w').write("\n".join(out))
