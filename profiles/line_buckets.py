import csv,re,sys
from collections import defaultdict
src_csv, dis_txt = sys.argv[1:3]
kname="macm_step_kernelILi32ELi2ELi0"; dname="macm_step_kernel<(int)32, (int)2, (int)0>"
lines=[];cur=None;infn=False
for ln in open(dis_txt,errors='replace'):
    if ln.startswith('.text.'):
        infn = kname in ln; continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=int(m.group(2)); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",ln)
    if m: lines.append((cur,m.group(2)))
rows=list(csv.reader(open(src_csv)))
start=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name' and dname in r[1]][0]
hdr=rows[start+1]
ci,cs,ct=hdr.index('Instructions Executed'),hdr.index('# Samples'),hdr.index('Thread Instructions Executed')
inst=[]
for r in rows[start+2:]:
    if not r or r[0]=='Kernel Name': break
    inst.append((r[1].strip(),int(r[ci] or 0),int(r[cs] or 0),int(r[ct] or 0)))
B=[(0,146,'grp prims'),(147,156,'bit helpers'),(157,163,'f32x2'),(164,183,'aabb/min/max/normalize'),(184,213,'wrap/atan2'),(214,236,'sincos'),(237,266,'solve_velocity'),(267,276,'warm_start'),(277,299,'solve_position'),(300,406,'find_new_contacts'),(407,508,'nn_search'),(509,519,'cartesian'),(520,581,'flock_observe'),(626,801,'big/fresh'),(802,885,'ph0 load'),(886,951,'ph1 actions'),(952,1027,'ph1b/2a'),(1028,1089,'ph2 collide'),(1090,1108,'ph3 integrate v'),(1109,1254,'ph4/5 islands+vel solver'),(1255,1277,'ph6 integrate x'),(1278,1323,'ph7 pos solver'),(1324,1356,'ph8 sleep'),(1357,1381,'ph9 sync fixtures'),(1382,1385,'ph10'),(1386,1443,'ph11/12 rewards+writeback'),(1444,1467,'ph13')]
ALU=('FSETP','ISETP','FSEL','SEL','VIADD','LOP3','IADD3','SHF','LEA','MOV','FMNMX','FMNMX3','PRMT','VIMNMX','VIMNMX3','PLOP3','IABS','ISCADD','BMSK','SGXT','FCHK','VOTE','P2R','R2P','CS2R','IMNMX')
FMA=('FADD','FMUL','FFMA','IMAD','FADD2','FMUL2','FFMA2','HFMA2','HADD2','HMUL2')
per=defaultdict(lambda:[0,0,0,0,0]);tot=[0,0,0,0,0]
n=min(len(inst),len(lines))
for k in range(n):
    ln=lines[k][0] or 0
    name='?'
    for lo,hi,nm in B:
        if lo<=ln<=hi: name=nm;break
    s=re.sub(r'^@!?U?P\d+\s+','',inst[k][0]); op=s.split()[0].split('.')[0]
    v=[inst[k][1],inst[k][2],inst[k][3], inst[k][1] if op in ALU else 0, inst[k][1] if op in FMA else 0]
    for j in range(5):
        per[name][j]+=v[j]; tot[j]+=v[j]
E=4096
print('total warp-inst %d (%.0f/env) samples %d thr/inst %.1f  ALU %.0f/env FMA %.0f/env'%(tot[0],tot[0]/E,tot[1],tot[2]/max(tot[0],1),tot[3]/E,tot[4]/E))
for nm,v in sorted(per.items(),key=lambda x:-x[1][0]):
    print('%-28s inst/env %6.0f %5.1f%%  samples %5.1f%%  thr/inst %4.1f  alu/env %5.0f fma/env %5.0f'%(nm,v[0]/E,100*v[0]/tot[0],100*v[1]/max(tot[1],1),v[2]/max(v[0],1),v[3]/E,v[4]/E))
