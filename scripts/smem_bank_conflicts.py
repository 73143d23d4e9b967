#!/usr/bin/env python3
"""Brute-force count of shared-memory wavefronts of the Fft3E exchange patterns
(fft_static.cuh: pass A/B stores, load_b, pass C loads) for candidate pitches PA / PB,
for a row CTA (one sequence) and a column CTA (C interleaved sequences), 8- and 16-byte
elements.  This is where the RA = 15 pitches (PA = 147, PB = 255) come from."""
# shared-memory wavefront count for the 8-byte-element exchange patterns of Fft3E
import itertools
def wavefronts(addrs, elem_bytes=8):
    # addrs: per-thread element index for one warp-wide access (None = inactive)
    # 8-byte accesses: processed per half-warp; count distinct 128B-bank rows conflicts
    total=0
    per = 128//elem_bytes  # threads per wavefront ideal
    for h in range(0,32,per):
        grp=[a for a in addrs[h:h+per] if a is not None]
        if not grp: continue
        # bank = (addr*elem_bytes/4) % 32 -> for 8B elems: bank pair index = addr % 16
        nb = 128//elem_bytes
        cnt={}
        for a in set(grp):
            cnt[a%nb]=cnt.get(a%nb,0)+1
        total+=max(cnt.values())
    return total
def plan(RA,RB,RC,NT,PA,PB):
    L=RA*RB*RC; NA=RB*RC; NB=RA*RC; NC=RA*RB
    MB=-(-NB//NT); MC=-(-NC//NT); MA=-(-NA//NT)
    ideal=0; actual=0
    nw=-(-NT//32)
    for w in range(nw):
        ts=[t if t<NT else None for t in range(32*w,32*w+32)]
        # pass A stores
        for m in range(MA):
            for q in range(RA):
                ad=[ (q*PA+t+m*NT) if (t is not None and t+m*NT<NA) else None for t in ts]
                if any(a is not None for a in ad):
                    actual+=wavefronts(ad); ideal+= -(-sum(a is not None for a in ad)//16)
        # load_b
        for m in range(MB):
            for q in range(RB):
                ad=[]
                for t in ts:
                    if t is None or t+m*NT>=NB: ad.append(None); continue
                    j=t+m*NT; k=j%RA; ad.append(k*PA+j//RA+q*RC)
                if any(a is not None for a in ad):
                    actual+=wavefronts(ad); ideal+= -(-sum(a is not None for a in ad)//16)
        # pass_b stores
        for m in range(MB):
            for q in range(RB):
                ad=[ (q*PB+t+m*NT) if (t is not None and t+m*NT<NB) else None for t in ts]
                if any(a is not None for a in ad):
                    actual+=wavefronts(ad); ideal+= -(-sum(a is not None for a in ad)//16)
        # pass_c loads
        for m in range(MC):
            for q in range(RC):
                ad=[]
                for t in ts:
                    if t is None or t+m*NT>=NC: ad.append(None); continue
                    j=t+m*NT; ad.append((j//RA)*PB+(j%RA)+q*RA)
                if any(a is not None for a in ad):
                    actual+=wavefronts(ad); ideal+= -(-sum(a is not None for a in ad)//16)
    return actual, ideal, max(RA*PA, RB*PB, L)
print('fwd current', plan(16,9,15,144,135,240))
print('inv current', plan(15,9,16,144,145,241))
best=[]
for PA in range(144,176):
    for PB in range(240,272):
        a,i,seq=plan(15,9,16,144,PA,PB)
        best.append((a,seq,PA,PB))
best.sort()
print(best[:12])
bestf=[]
for PA in range(135,160):
    for PB in range(240,260):
        a,i,seq=plan(16,9,15,144,PA,PB)
        bestf.append((a,seq,PA,PB))
bestf.sort(); print(bestf[:8])
print('--- inverse candidates with SEQ<=2304')
c=[b for b in best if b[1]<=2304]
print(c[:10])
print('--- 16-byte elements (fp64)')
def plan16(RA,RB,RC,NT,PA,PB):
    global wavefronts
    import types
    L=RA*RB*RC; NA=RB*RC; NB=RA*RC; NC=RA*RB
    MB=-(-NB//NT); MC=-(-NC//NT); MA=-(-NA//NT)
    actual=0; ideal=0
    for w in range(-(-NT//32)):
        ts=[t if t<NT else None for t in range(32*w,32*w+32)]
        def acc(ad):
            nonlocal actual, ideal
            if any(a is not None for a in ad):
                actual+=wavefronts(ad,16); ideal+=-(-sum(a is not None for a in ad)//8)
        for m in range(MB):
            for q in range(RB):
                ad=[]
                for t in ts:
                    if t is None or t+m*NT>=NB: ad.append(None); continue
                    j=t+m*NT; k=j%RA; ad.append(k*PA+j//RA+q*RC)
                acc(ad)
        for m in range(MC):
            for q in range(RC):
                ad=[]
                for t in ts:
                    if t is None or t+m*NT>=NC: ad.append(None); continue
                    j=t+m*NT; ad.append((j//RA)*PB+(j%RA)+q*RA)
                acc(ad)
    return actual, ideal
print('inv fp64 current', plan16(15,9,16,144,145,241), 'new', plan16(15,9,16,144,146,255), 'alt', plan16(15,9,16,144,145,255))
print('fwd fp64', plan16(16,9,15,144,135,240))
print('--- column CTA (C interleaved sequences, tid = t*C + c, sequence c at c*LSM)')
def colplan(RA,RB,RC,NT,PA,PB,C=4):
    NA=RB*RC; NB=RA*RC; NC=RA*RB
    MB=-(-NB//NT); MC=-(-NC//NT); MA=-(-NA//NT)
    SEQ=max(RA*PA,RB*PB,RA*RB*RC); LSM=(SEQ+15)//16*16+16//C
    actual=0; ideal=0
    for w in range(NT*C//32):
        tids=range(32*w,32*w+32)
        def acc(f):
            nonlocal actual, ideal
            ad=[]
            for tid in tids:
                t=tid//C; c=tid%C
                p=f(t)
                ad.append(None if p is None else c*LSM+p)
            if any(a is not None for a in ad):
                actual+=wavefronts(ad); ideal+=-(-sum(a is not None for a in ad)//16)
        for m in range(MA):
            for q in range(RA):
                acc(lambda t: (q*PA+t+m*NT) if t+m*NT<NA else None)
        for m in range(MB):
            for q in range(RB):
                acc(lambda t: ((t+m*NT)%RA*PA+(t+m*NT)//RA+q*RC) if t+m*NT<NB else None)
        for m in range(MB):
            for q in range(RB):
                acc(lambda t: (q*PB+t+m*NT) if t+m*NT<NB else None)
        for m in range(MC):
            for q in range(RC):
                acc(lambda t: ((t+m*NT)//RA*PB+(t+m*NT)%RA+q*RA) if t+m*NT<NC else None)
    return actual, ideal, LSM
for PA,PB in ((145,241),(146,255),(145,255),(147,255),(149,255),(159,255)):
    print('inv', PA, PB, colplan(15,9,16,144,PA,PB))
print('fwd', colplan(16,9,15,144,135,240))
print('fp64 rows inv (147,255)', plan16(15,9,16,144,147,255), '(145,255)', plan16(15,9,16,144,145,255))
print('fp32 rows inv (147,255)', plan(15,9,16,144,147,255))
print('--- row plan 12 x 12 x 15, 180 threads')
for PA in (180, 181, 183, 185, 187, 189, 191):
    for PB in (180, 181, 183, 185, 187, 189, 191):
        a, i, seq = plan(12, 12, 15, 180, PA, PB)
        if a <= i * 1.02: print('fwd', PA, PB, a, i, seq)
best = []
for PA in range(144, 160):
    for PB in range(180, 200):
        a, i, seq = plan(15, 12, 12, 180, PA, PB)
        best.append((a, seq, PA, PB, i))
best.sort(); print('inv', best[:6])
bestf = []
for PA in range(180, 200):
    for PB in range(180, 200):
        a, i, seq = plan(12, 12, 15, 180, PA, PB)
        bestf.append((a, seq, PA, PB, i))
bestf.sort(); print('fwd', bestf[:6])
