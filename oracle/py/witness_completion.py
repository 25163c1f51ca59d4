"""Witness completion by constraint propagation -- TEST INFRASTRUCTURE (oracle side), not part of the product.

The reference's withdraw circuit (/root/reference/noir_circuit/target/shielded_pool_verifier.ccs) declares every ACIR
witness its constraints read -- 6,184 of them -- as a secret input, and the reference commits no witness file
(`nargo execute` output is git-ignored).  This module rebuilds a full, satisfying assignment from the circuit's ABI
inputs alone (/root/reference/noir_circuit/src/main.nr:39-53, values e.g. /root/reference/client/prover-params.toml),
WITHOUT an ACVM: the R1CS itself determines the intermediate witnesses.

  1. ordinary propagation: a row with one unknown wire defines it (as gnark's solver does, but in data-flow order and
     for "input" wires too); a hint runs once its inputs are known;
  2. radix rows: a linear row whose unknowns carry distinct powers of two is a limb / bit decomposition written by an
     unconstrained (Brillig) function -- split the known side accordingly, widths from the range checks;
  3. small linear systems: what is left are the is-zero / comparison gadgets (x * inv = 1 - z, x * z = 0); rows that are
     linear in their unknowns are solved by Gaussian elimination per connected component; a variable that stays free
     (the inverse witness of a zero) is set to 0.

Every result is checked row by row by the caller (tests/test_withdraw_real.py) and by the strict gnark-order solver
(groth16.solve).  Given only the PRIVATE inputs of prover-params.toml it reproduces that file's public inputs (root,
nullifier, wa_commitment) and the public key (owner_x, owner_y) -- which pins the solver, the Poseidon constraints as
parsed, and the three sw-grumpkin / emulated hints on the reference's own test vector."""
import collections

import groth16 as G

R = G.R
CONST = 0xFFFFFFFF


def inv(x):
    return pow(x, R - 2, R)


def complete(c, abi_values, pk=None, blinder=7, verbose=False):
    """abi_values: {wire: value} for the known input wires.  -> (wires [None = unknown], unknown wire list, violated
    check rows, stuck rows, instruction table, commitments).  Without `pk` the BSB22 commitment hint is skipped: the
    input wires (the assignment) are complete, the gnark-internal wires behind the commitment stay unknown."""
    nw=c.nb_wires
    w=[None]*nw; w[0]=1
    for k,v in abi_values.items(): w[k]=v%R
    names=c.body['MHintsDependencies']
    rows=[]   # (kind, data)
    # R1C rows
    instr_rows=[]
    for ins,bp in enumerate(c.blueprint):
        if bp==1:
            L,Rr,O=c.r1c(ins); instr_rows.append(('r1c',ins,(L,Rr,O)))
        else:
            hid,iex,o0,o1=c.hint(ins); instr_rows.append(('hint',ins,(names[hid],iex,o0,o1)))
    # wire -> rows using it
    uses=collections.defaultdict(list)
    for idx,(kind,ins,d) in enumerate(instr_rows):
        ws=set()
        if kind=='r1c':
            for e in d:
                for cid,wid in e:
                    if wid!=CONST: ws.add(wid)
        else:
            for e in d[1]:
                for cid,wid in e:
                    if wid!=CONST: ws.add(wid)
        for x in ws: uses[x].append(idx)
        instr_rows[idx]=(kind,ins,d,ws)
    # range info: wire -> bits
    rng_bits={}
    for kind,ins,d,ws in instr_rows:
        if kind=='hint':
            name,iex,o0,o1=d
            if name.endswith('rangecheck.DecomposeHint'):
                varsize=c.coeffs[iex[0][0][0]]; 
                e=iex[2]
                if len(e)==1 and c.coeffs[e[0][0]]==1 and e[0][1]!=CONST:
                    rng_bits[e[0][1]]=min(rng_bits.get(e[0][1],999), varsize)
            if name.endswith('bits.nBits'):
                e=iex[0]
                if len(e)==1 and c.coeffs[e[0][0]]==1: rng_bits[e[0][1]]=min(rng_bits.get(e[0][1],999), o1-o0)
    # booleans: rows b*(1-b)=0
    for kind,ins,d,ws in instr_rows:
        if kind=='r1c':
            L,Rr,O=d
            if len(L)==1 and len(ws)==1 and len(Rr)==2 and len(O)<=1:
                b=L[0][1]
                if b!=CONST and c.coeffs[L[0][0]]==1:
                    rr={wid:c.coeffs[cid] for cid,wid in Rr}
                    if rr.get(0)==1 and rr.get(b)==R-1 and all(c.coeffs[cid]==0 for cid,wid in O):
                        rng_bits[b]=min(rng_bits.get(b,999),1)
    def lin(e):
        s=0; unk={}
        for cid,wid in e:
            co=c.coeffs[cid]
            if wid==CONST: s+=co
            elif w[wid] is None: unk[wid]=(unk.get(wid,0)+co)%R
            else: s+=co*w[wid]
        return s%R, {k:v for k,v in unk.items() if v}
    done=[False]*len(instr_rows)
    queue=collections.deque(range(len(instr_rows)))
    inq=[True]*len(instr_rows)
    commitments=[]
    failed=[]
    def setw(x,v):
        if w[x] is not None: return
        w[x]=v%R
        for idx in uses[x]:
            if not done[idx] and not inq[idx]:
                queue.append(idx); inq[idx]=True
    stuck_lin=[]
    queue_items=[]
    def propagate():
      while queue:
          idx=queue.popleft(); inq[idx]=False
          if done[idx]: continue
          kind,ins,d,ws=instr_rows[idx]
          if kind=='hint':
              name,iex,o0,o1=d
              if any(w[x] is None for x in ws): continue
              vals=[lin(e)[0] for e in iex]
              if name==G.HINT_RANDOMIZE: outs=[blinder]*(o1-o0)
              elif name==G.HINT_COMMIT:
                  if pk is None:
                      done[idx]=True      # the assignment (input wires) does not depend on it
                      continue
                  info=c.commitments[len(commitments)]
                  npc=len(info["PublicAndCommitmentCommitted"])
                  committed=vals[1+npc:]
                  key=pk["commitment_keys"][len(commitments)]
                  import bn254 as B, serialize as S
                  pt=B.g1_msm(key["basis"], committed)
                  msg=S.g1_to_bytes(pt)+b"".join(S.fr_to_bytes(x) for x in vals[1:1+npc])
                  outs=[G.hash_to_field(msg, G.COMMITMENT_DST)[0]]
                  commitments.append((pt,committed))
              else: outs=G.HINTS[name](vals,o1-o0)
              for k,v in enumerate(outs): setw(o0+k,v)
              done[idx]=True
              continue
          L,Rr,O=d
          a,ua=lin(L); b,ub=lin(Rr); cc,uc=lin(O)
          unk=set(ua)|set(ub)|set(uc)
          if not unk:
              done[idx]=True
              if a*b%R!=cc: failed.append(c.constraint_offset[ins])
              continue
          if len(unk)==1:
              x=next(iter(unk))
              ina,inb,inc=x in ua,x in ub,x in uc
              if inc and not ina and not inb:
                  setw(x,(a*b-cc)*inv(uc[x])); done[idx]=True; continue
              if ina and not inb and not inc:
                  if b==0:
                      # a*0 = c : x free; gnark sets 0
                      setw(x,0)
                  else: setw(x,(cc*inv(b)-a)*inv(ua[x]))
                  done[idx]=True; continue
              if inb and not ina and not inc:
                  if a==0: setw(x,0)
                  else: setw(x,(cc*inv(a)-b)*inv(ub[x]))
                  done[idx]=True; continue
              if ina and inc and not inb:
                  # (a + ca x) b = cc + cc_x x  -> x (ca b - ccx) = cc - a b
                  den=(ua[x]*b-uc[x])%R
                  if den: setw(x,(cc-a*b)*inv(den)); done[idx]=True
                  continue
              if inb and inc and not ina:
                  den=(ub[x]*a-uc[x])%R
                  if den: setw(x,(cc-a*b)*inv(den)); done[idx]=True
                  continue
              continue
          # several unknowns: linear rows only
          if not ua and not ub:
              # a*b = cc + sum uc  -> sum c_i u_i = a*b - cc
              K=(a*b-cc)%R; coefs=dict(uc)
          elif not ub and not uc and not ua=={} and False:
              continue
          elif not ua and not uc:
              # a*(b+sum ub)=cc -> sum = cc/a - b
              if a==0: continue
              K=(cc*inv(a)-b)%R; coefs=dict(ub)
          elif not ub and not uc:
              if b==0: continue
              K=(cc*inv(b)-a)%R; coefs=dict(ua)
          else:
              continue
          # decomposition rule
          items=[]
          ok=True
          sign=None
          for x,co in coefs.items():
              neg = co>R//2
              m = R-co if neg else co
              if m & (m-1): ok=False; break
              if sign is None: sign=neg
              elif sign!=neg: ok=False; break
              items.append((m.bit_length()-1,x))
          if not ok:
              stuck_lin.append((idx,len(coefs))); continue
          Kp=(R-K)%R if sign else K
          items.sort()
          if any(items[i][0]==items[i+1][0] for i in range(len(items)-1)):
              stuck_lin.append((idx,len(coefs))); continue
          vals={}
          good=True
          for i,(e,x) in enumerate(items):
              nxt=items[i+1][0] if i+1<len(items) else None
              bw=rng_bits.get(x)
              width = (nxt-e) if nxt is not None else None
              if bw is not None and width is not None: width_use=min(bw,width)
              elif bw is not None: width_use=bw
              else: width_use=width
              v=(Kp>>e) if width_use is None else (Kp>>e)&((1<<width_use)-1)
              vals[x]=v
          if not good: stuck_lin.append((idx,len(coefs))); continue
          if sum(v<<e for (e,x),v in zip(items,[vals[x] for e,x in items]))!=Kp:
              stuck_lin.append((idx,-len(coefs))); continue
          for x,v in vals.items(): setw(x,v)
          done[idx]=True
    propagate()
    # ---- stalled: small linear systems (is-zero / comparison gadgets written by Brillig) -------------------
    progress=True
    rounds=0
    fallback=False
    fallback_tried=False
    while progress or not fallback_tried:
        if not progress:
            fallback=True; fallback_tried=True
        else:
            fallback_tried=False
        progress=False
        rounds+=1
        eqs=[]   # (dict unknown->coef, rhs)
        for idx,(kind,ins,d,ws) in enumerate(instr_rows):
            if done[idx] or kind!='r1c': continue
            L,Rr,O=d
            a,ua=lin(L); b,ub=lin(Rr); cc,uc=lin(O)
            if ua and ub: continue
            # (a+ua)(b) = cc+uc  or a (b+ub) = cc+uc
            co={}
            if ua:
                for x,v in ua.items(): co[x]=(co.get(x,0)+v*b)%R
            if ub:
                for x,v in ub.items(): co[x]=(co.get(x,0)+v*a)%R
            for x,v in uc.items(): co[x]=(co.get(x,0)-v)%R
            co={x:v for x,v in co.items() if v}
            rhs=(cc-a*b)%R
            if 0<len(co)<=4: eqs.append((co,rhs,idx))
        if not eqs: break
        # connected components over unknowns
        parent={}
        def find(x):
            while parent.setdefault(x,x)!=x:
                parent[x]=parent[parent[x]]; x=parent[x]
            return x
        for co,rhs,idx in eqs:
            xs=list(co)
            for y in xs[1:]: parent[find(y)]=find(xs[0])
        comps=collections.defaultdict(list)
        for e in eqs: comps[find(next(iter(e[0])))].append(e)
        for root,es in comps.items():
            xs=sorted({x for co,_,_ in es for x in co})
            if len(xs)>40: continue
            col={x:i for i,x in enumerate(xs)}
            M=[[0]*(len(xs)+1) for _ in es]
            for r,(co,rhs,idx) in enumerate(es):
                for x,v in co.items(): M[r][col[x]]=v
                M[r][-1]=rhs
            # gaussian elimination
            piv=[]; rr=0
            for cidx in range(len(xs)):
                p=None
                for r in range(rr,len(M)):
                    if M[r][cidx]: p=r; break
                if p is None: continue
                M[rr],M[p]=M[p],M[rr]
                iv=inv(M[rr][cidx])
                M[rr]=[v*iv%R for v in M[rr]]
                for r in range(len(M)):
                    if r!=rr and M[r][cidx]:
                        f=M[r][cidx]
                        M[r]=[(v-f*u)%R for v,u in zip(M[r],M[rr])]
                piv.append(cidx); rr+=1
            # inconsistent?
            if any(all(v==0 for v in row[:-1]) and row[-1] for row in M): continue
            if len(piv)<len(xs):
                # free variables -> 0 (only if the system is otherwise consistent)
                pass
            free=[i for i in range(len(xs)) if i not in piv]
            for r,cidx in enumerate(piv):
                # determined only if the pivot row does not involve a free variable (unless we are in fallback mode)
                if fallback or all(M[r][f]==0 for f in free):
                    x=xs[cidx]
                    if w[x] is None: setw(x,M[r][-1]); progress=True
            if fallback:
                for f in free:
                    if w[xs[f]] is None: setw(xs[f],0); progress=True
                if progress: fallback=False; break
        # continue ordinary propagation
        while queue:
            idx=queue.popleft(); inq[idx]=False
            if done[idx]: continue
            queue_items.append(idx)
        if queue_items:
            pending=list(queue_items); queue_items.clear()
            for idx in pending: queue.append(idx); inq[idx]=True
            propagate()
    used=set(uses)
    for i in range(nw):
        if w[i] is None and i not in used and not any(False for _ in ()): 
            # a wire no instruction reads or defines is free
            w[i]=0
    unknown=[i for i,v in enumerate(w) if v is None]
    if verbose:
        print("unknown wires:",len(unknown),"of",nw,"failed checks:",len(failed),failed[:10], "stuck linear rows:",len(stuck_lin))
    return w,unknown,failed,stuck_lin,instr_rows,commitments

