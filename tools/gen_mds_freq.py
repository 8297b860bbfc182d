# derive the frequency-domain constants of the 12x12 circulant MDS (4-point DFT over stride-3 subsequences)
C = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
Cp = [C[(12 - j) % 12] for j in range(12)]   # out = s (*) Cp cyclic
from fractions import Fraction as F
def K(b, y0):  # y0 in {1,-1,1j}
    return sum(Cp[3*a+b] * (y0 ** a) for a in range(4))
G1 = [[None]*3 for _ in range(3)]; Gm = [[None]*3 for _ in range(3)]; Gi = [[None]*3 for _ in range(3)]
for c in range(3):
    for b in range(3):
        bp = (c - b) % 3
        wrap = b + bp >= 3
        G1[c][b] = F(K(bp, 1) * (1 if wrap else 1), 4)
        Gm[c][b] = F(K(bp, -1) * (-1 if wrap else 1), 4)
        z = K(bp, 1j) * (1j if wrap else 1)
        Gi[c][b] = (F(int(z.real), 2), F(int(z.imag), 2))
print("G1", [[float(x) for x in r] for r in G1])
print("Gm", [[float(x) for x in r] for r in Gm])
print("Gi", [[(float(x), float(y)) for x, y in r] for r in Gi])
import random
def mds_ref(s):
    return [sum(C[i] * s[(i + r) % 12] for i in range(12)) + (8 * s[0] if r == 0 else 0) for r in range(12)]
def mds_freq(s):
    S1=[0]*3;Sm=[0]*3;Sr=[0]*3;Si=[0]*3
    for b in range(3):
        e0=s[b]+s[6+b]; e1=s[3+b]+s[9+b]; d0=s[b]-s[6+b]; d1=s[3+b]-s[9+b]
        S1[b]=e0+e1; Sm[b]=e0-e1; Sr[b]=d0; Si[b]=d1
    out=[0]*12
    for c in range(3):
        A=sum(S1[b]*G1[c][b] for b in range(3)); B=sum(Sm[b]*Gm[c][b] for b in range(3))
        P=sum(Sr[b]*Gi[c][b][0]-Si[b]*Gi[c][b][1] for b in range(3)); Q=sum(Sr[b]*Gi[c][b][1]+Si[b]*Gi[c][b][0] for b in range(3))
        u=A+B; v=A-B
        out[c]=u+P; out[3+c]=v+Q; out[6+c]=u-P; out[9+c]=v-Q
    out[0]+=8*s[0]
    return out
for _ in range(100):
    s=[random.randrange(2**32) for _ in range(12)]
    assert mds_ref(s)==[int(x) for x in mds_freq(s)], (mds_ref(s), mds_freq(s))
print("ok")
