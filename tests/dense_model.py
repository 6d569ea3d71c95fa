"""NumPy model of the *plane* (tile-sharing) formulation that ssl_b200/csrc/ssg_plane.cuh implements.

Test infrastructure (imported by tests/test_dense_model.py only).  It spells out, per search
offset d = (dy, dx), the algebra the dense sm_100a kernels use instead of the per-edge-pixel loops
of similarity.cu:5-131, so that the index arithmetic can be checked against the oracle on the CPU:

  forward   D_d(Y,X)   = sum_c (I(Y,X) - I(Y+dy,X+dx))^2                     (padded coordinates)
            Sh_d(Y,Xs) = sum_{k<l(dx)} D_d(Y, Xs-k)                          "sum of the last l"
            q(p,d)     = sum_{a in A(dy)} Sh_d(py+a, px+hi(dx)) + Eout(p,d)
  backward  u_d        = sparse plane holding gq(p,d) and gq(p,-d) (placement rules below)
            Gs_d(Y,X)  = sum_{k<l(dx)} u_d(Y, X+K-k)
            dL/dIpad(Y,X,c) = 2 sum_d (I(Y,X,c) - I(Y+dy,X+dx,c)) Gs_d(Y,X) + 2 I(Y,X,c) Wout(Y,X)

with A(t) = [lo(t), hi(t)] = [max(-K,-P-t), min(K,P-t)] the in-area part of the window for offset
t (the zero-padded unfold of loss_util.py:208-209 / the bounds test of similarity.cu:43), l(t) its
length, and Eout / Wout the out-of-area terms (neighbour treated as zero, similarity.cu:46-47,123-124).
"""
import numpy as np


def lo(t, P, K):
    return max(-K, -P - t)


def hi(t, P, K):
    return min(K, P - t)


def reflect_pad(img, P):
    return np.pad(img, ((0, 0), (P, P), (P, P)), mode="reflect")


def forward_planes(img, pos, ks, kw):
    """q[mc, ks*ks] via displacement planes; img [C,H,W] float64, pos [mc,2] unpadded (y,x)."""
    P, K = ks // 2, kw // 2
    ip = reflect_pad(img, P)
    C, Hp, Wp = ip.shape
    big = np.zeros((C, Hp + 2 * P, Wp + 2 * P), ip.dtype)      # zero outside the padded image
    big[:, P:P + Hp, P:P + Wp] = ip
    E = (ip ** 2).sum(0)
    q = np.zeros((len(pos), ks * ks), ip.dtype)
    for dy in range(-P, P + 1):
        for dx in range(-P, P + 1):
            nb = big[:, P + dy:P + dy + Hp, P + dx:P + dx + Wp]
            D = ((ip - nb) ** 2).sum(0)
            l = hi(dx, P, K) - lo(dx, P, K) + 1
            Sh = np.zeros_like(D)
            for k in range(l):
                Sh[:, k:] += D[:, :Wp - k]
            for n, (y, x) in enumerate(pos):
                py, px = y + P, x + P
                acc = 0.0
                for a in range(lo(dy, P, K), hi(dy, P, K) + 1):
                    acc += Sh[py + a, px + hi(dx, P, K)]
                for a in range(-K, K + 1):
                    for b in range(-K, K + 1):
                        inside = lo(dy, P, K) <= a <= hi(dy, P, K) and lo(dx, P, K) <= b <= hi(dx, P, K)
                        if not inside:
                            acc += E[py + a, px + b]
                q[n, (dy + P) * ks + dx + P] = acc
    return q


def backward_planes(img, pos, gq, ks, kw):
    """dL/dimg [C,H,W] from gq = dL/dq [mc, ks*ks] via combined planes + reflect-pad adjoint."""
    P, K = ks // 2, kw // 2
    ip = reflect_pad(img, P)
    C, Hp, Wp = ip.shape
    M = P + 2 * K + 2   # margin so every placement lands inside the working plane
    big = np.zeros((C, Hp + 2 * P, Wp + 2 * P), ip.dtype)
    big[:, P:P + Hp, P:P + Wp] = ip
    gpad = np.zeros_like(ip)
    wout = np.zeros((Hp, Wp), ip.dtype)
    for dy in range(-P, P + 1):
        for dx in range(-P, P + 1):
            d = (dy + P) * ks + dx + P
            dm = (-dy + P) * ks + (-dx + P)
            u = np.zeros((Hp + 2 * M, Wp + 2 * M), ip.dtype)
            alo, ahi = lo(dy, P, K), hi(dy, P, K)
            blo, bhi = lo(dx, P, K), hi(dx, P, K)
            l = bhi - blo + 1
            for n, (y, x) in enumerate(pos):
                py, px = y + P, x + P
                # first kind: gq(p, d) at z = p; rows z+A, column px + K + blo
                for a in range(alo, ahi + 1):
                    u[M + py + a, M + px + K + blo] += gq[n, d]
                # second kind: gq(p, -d) at z = p - d; rows z + (-A), column zx + K - bhi
                for a in range(-ahi, -alo + 1):
                    u[M + py - dy + a, M + px - dx + K - bhi] += gq[n, dm]
                # out-of-area weights of (p, d)
                for a in range(-K, K + 1):
                    for b in range(-K, K + 1):
                        if not (alo <= a <= ahi and blo <= b <= bhi):
                            wout[py + a, px + b] += gq[n, d]
            Sh = np.zeros_like(u)
            for k in range(l):
                Sh[:, k:] += u[:, :u.shape[1] - k]
            Gs = Sh[M:M + Hp, M + K:M + K + Wp]
            nb = big[:, P + dy:P + dy + Hp, P + dx:P + dx + Wp]
            gpad += 2.0 * (ip - nb) * Gs[None]
    gpad += 2.0 * ip * wout[None]
    # adjoint of F.pad(reflect) (similaritywrapper.py:64)
    H, W = img.shape[1:]
    g = np.zeros_like(img)
    for Y in range(Hp):
        v = Y - P
        sy = -v if v < 0 else (2 * (H - 1) - v if v > H - 1 else v)
        for X in range(Wp):
            w_ = X - P
            sx = -w_ if w_ < 0 else (2 * (W - 1) - w_ if w_ > W - 1 else w_)
            g[:, sy, sx] += gpad[:, Y, X]
    return g
