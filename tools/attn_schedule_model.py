import heapq, math
def items_c(B, L, heads=32):
    nqp = math.ceil(L/256); n_groups=B*heads
    wave_groups=max(1, math.ceil(148/nqp))
    per_wave=wave_groups*nqp
    out=[]
    for item in range(nqp*n_groups):
        wave=item//per_wave; r=item%per_wave
        gw=min(wave_groups, n_groups-wave*wave_groups)
        g=wave*wave_groups + r%gw; qp=nqp-1-r//gw
        q0=qp*256
        if q0>=L: out.append(0); continue
        nA=2*qp+1; nB=2*qp+2 if q0+128<L else 0
        out.append(max(nA,nB))
    return out
def sim(items, blk=3150, ovh=3500, ncta=148):
    h=[0]*ncta; heapq.heapify(h)
    for n in items:
        t=heapq.heappop(h)
        heapq.heappush(h, t + (n*blk+ovh if n else 200))
    return max(h), sum(h)/ncta, min(h)
for name,B,L in (("c2",8,1485),("c4x2",2,2564),("c3 64",64,1485)):
    it=items_c(B,L)
    mx,avg,mn=sim(it)
    ideal=sum((n*3150+3500) for n in it if n)/148
    print(name, "finish",mx,"ideal",round(ideal),"eff",round(ideal/mx,3), "min",mn)
    # alternative: last two waves globally heaviest first
    nqp=math.ceil(L/256); wg=max(1,math.ceil(148/nqp)); pw=wg*nqp
    nw=math.ceil(len(it)/pw)
    for k in (1,2,3):
        cut=max(0,(nw-k))*pw
        alt=it[:cut]+sorted(it[cut:],reverse=True)
        mx2,_,_=sim(alt); print("   last",k,"waves LPT: finish",mx2,"eff",round(ideal/mx2,3))
    mx3,_,_=sim(sorted(it,reverse=True)); print("   global LPT", mx3, round(ideal/mx3,3))
print("K9")
for name,B,L in (("c2",8,1485),("c4x2",2,2564)):
    nb=math.ceil(L/128)
    it=[]
    for g in range(B*32):
        it += list(range(nb,0,-1))
    for blk,ovh in ((1000,3000),):
        mx,avg,mn=sim(it,blk,ovh); ideal=sum(n*blk+ovh for n in it)/148
        print(name,"group-major heavy-first: eff",round(ideal/mx,3))
        # waves of 148//nb groups, heaviest first in wave; last 3 merged
        mx2,_,_=sim(sorted(it,reverse=True),blk,ovh); print("   global LPT eff",round(ideal/mx2,3))
