import numpy as np
rng = np.random.default_rng(0)
N=50; L=np.array([2.0,2.3,1.9]); K=np.array([5,7,5]); alpha=3.0; ke=138.935456
pos = rng.uniform(-1,3,size=(N,3)); q = rng.normal(size=N); q-=q.mean()
C = 4*np.pi*ke/np.prod(L)
# brute force (reference loops)
E=0; F=np.zeros((N,3)); dedq=np.zeros(N)
for nx in range(K[0]):
  for ny in range(0 if nx==0 else 1-K[1], K[1]):
    for nz in range(1 if (nx==0 and ny==0) else 1-K[2], K[2]):
      k = 2*np.pi*np.array([nx,ny,nz])/L; k2=k@k; ak=np.exp(-k2/(4*alpha**2))/k2
      gr = pos@k; cs=(q*np.cos(gr)).sum(); ss=(q*np.sin(gr)).sum()
      E += C*ak*(cs*cs+ss*ss)
      g = 2*C*ak*(ss*q*np.cos(gr)-cs*q*np.sin(gr)); F -= g[:,None]*k[None,:]
      dedq += 2*C*ak*(cs*np.cos(gr)+ss*np.sin(gr))
# factorised
u = pos/L; u -= np.floor(u)
tab = [np.exp(2j*np.pi*np.arange(K[a])[:,None]*u[:,a][None,:]) for a in range(3)]  # [n][atom]
Ex,Ey,Ez = tab
X = q[None,:]*Ex
# P sums: rows (nx,m), cols l
a0 = np.einsum('xj,mj->xmj', X.real, Ey.real); a1=np.einsum('xj,mj->xmj', X.real, Ey.imag)
a2 = np.einsum('xj,mj->xmj', X.imag, Ey.real); a3=np.einsum('xj,mj->xmj', X.imag, Ey.imag)
zc, zs = Ez.real, Ez.imag
P = {}
for name,a in (('rc',a0),('rs',a1),('ic',a2),('is',a3)):
    P[name+'c'] = np.einsum('xmj,lj->xml', a, zc); P[name+'s']=np.einsum('xmj,lj->xml', a, zs)
def S(nx,ny,nz):
    m,l=abs(ny),abs(nz); sy=1 if ny>=0 else -1; sz=1 if nz>=0 else -1
    re = P['rcc'][nx,m,l]-sy*sz*P['rss'][nx,m,l]-sz*P['ics'][nx,m,l]-sy*P['isc'][nx,m,l]
    im = P['icc'][nx,m,l]-sy*sz*P['iss'][nx,m,l]+sz*P['rcs'][nx,m,l]+sy*P['rsc'][nx,m,l]
    return re+1j*im
def included(nx,ny,nz):
    if nx>0: return True
    if ny>0: return True
    return ny==0 and nz>0
E2=0
G={}
for nx in range(K[0]):
  for ny in range(1-K[1],K[1]):
    for nz in range(1-K[2],K[2]):
      if not included(nx,ny,nz): G[(nx,ny,nz)]=0; continue
      k = 2*np.pi*np.array([nx,ny,nz])/L; k2=k@k; ak=np.exp(-k2/(4*alpha**2))/k2
      s=S(nx,ny,nz); E2+=C*ak*abs(s)**2; G[(nx,ny,nz)]=2*C*ak*s
print("E", E, E2)
F2=np.zeros((N,3)); d2=np.zeros(N)
for nx in range(K[0]):
  for ny in range(1-K[1],K[1]):
    if nx==0 and ny<0: continue
    U=np.zeros(N,complex); Up=np.zeros(N,complex)
    for l in range(K[2]):
      if l==0:
        A=np.conj(G[(nx,ny,0)]); B=0
      else:
        Hp=np.conj(G[(nx,ny,l)]); Hm=np.conj(G[(nx,ny,-l)]); A=Hp+Hm; B=1j*(Hp-Hm)
      U += A*zc[l]+B*zs[l]
      Up += l*((-1j*B)*zc[l]+(1j*A)*zs[l])
    T = Ex[nx]*(Ey[ny] if ny>=0 else np.conj(Ey[-ny]))
    TU=T*U; TUp=T*Up
    d2 += TU.real
    F2[:,0]+= q*(2*np.pi/L[0])*nx*TU.imag
    F2[:,1]+= q*(2*np.pi/L[1])*ny*TU.imag
    F2[:,2]+= q*(2*np.pi/L[2])*TUp.imag
print("F", np.abs(F-F2).max(), np.abs(F).max())
print("dedq", np.abs(dedq-d2).max(), np.abs(dedq).max())
