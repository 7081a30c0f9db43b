# PTX micro-benchmarks of the FP64 pipe with the exact operand patterns of the PIP leaf code
def kernel(name, body_fn, N=16):
    L=[]
    w=L.append
    w(f".visible .entry {name}(.param .u64 p_out, .param .u32 p_iters)")
    w(".maxntid 256, 1, 1")
    w("{")
    w(f".reg .f64 %c<{N}>, %x<{N}>, %g<{N}>, %v<{N}>, %a, %s; .reg .u64 %po, %off; .reg .u32 %it, %n, %t, %r<3>; .reg .pred %p;")
    w("ld.param.u64 %po, [p_out]; ld.param.u32 %n, [p_iters]; cvta.to.global.u64 %po, %po;")
    w("mov.u32 %r0, %ctaid.x; mov.u32 %r1, %ntid.x; mov.u32 %r2, %tid.x; mad.lo.u32 %t, %r0, %r1, %r2;")
    w("cvt.rn.f64.u32 %s, %t; mul.rn.f64 %s, %s, 0d3E10000000000000;")   # tiny thread-dependent value
    for i in range(N):
        w(f"add.rn.f64 %c{i}, %s, 0d{0x3FEFFFF000000000+i*4096:016X}; add.rn.f64 %x{i}, %s, 0d{0x3FEFFFE000000000+i*8192:016X};")
        w(f"mov.f64 %g{i}, %s; mov.f64 %v{i}, %s;")
    w("add.rn.f64 %a, %s, 0d3FEFFFFF00000000;")
    w("mov.u32 %it, 0;")
    w("LOOP:")
    body_fn(w, N)
    w("add.u32 %it, %it, 1; setp.lt.u32 %p, %it, %n; @%p bra LOOP;")
    w("mov.f64 %s, %a;")
    for i in range(N):
        w(f"add.rn.f64 %s, %s, %g{i}; add.rn.f64 %s, %s, %v{i};")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off; st.global.f64 [%po], %s;")
    w("ret;")
    w("}")
    return L
def leaf(nacc):
    def f(w,N):
        for rep in range(2):
            for i in range(N):
                w(f"fma.rn.f64 %v{i%nacc}, %c{i}, %x{i}, %v{i%nacc};")
                w(f"fma.rn.f64 %g{i}, %c{i}, %a, %g{i};")
    return f
def grouped(nacc):
    def f(w,N):
        for rep in range(2):
            for i in range(N): w(f"fma.rn.f64 %v{i%nacc}, %c{i}, %x{i}, %v{i%nacc};")
            for i in range(N): w(f"fma.rn.f64 %g{i}, %c{i}, %a, %g{i};")
    return f
def gonly(w,N):
    for rep in range(4):
        for i in range(N): w(f"fma.rn.f64 %g{i}, %c{i}, %a, %g{i};")
def vonly(nacc):
    def f(w,N):
        for rep in range(4):
            for i in range(N): w(f"fma.rn.f64 %v{i%nacc}, %c{i}, %x{(i+rep)%N}, %v{i%nacc};")
    return f
ks={'leaf2':leaf(2),'leaf4':leaf(4),'leaf8':leaf(8),'leaf16':leaf(16),'grouped2':grouped(2),'grouped4':grouped(4),'gonly':gonly,'vonly2':vonly(2),'vonly4':vonly(4),'vonly16':vonly(16)}
out=[".version 8.7",".target sm_100a",".address_size 64",""]
for n,f in ks.items(): out+=kernel('ub_'+n,f)+[""]
open('build/ubench/ub.ptx','w').write("\n".join(out))
print(list(ks))
