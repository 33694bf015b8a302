"""Two implementations behind one numpy-in / numpy-out interface, so that the same cases check both:

  OracleImpl  — the CPU oracle (oracle/oracle.c), used without a GPU to pin the oracle to the goldens;
  CudaImpl    — the PRODUCT: libmvp_ops.so called through its C ABI (ctypes, raw device pointers;
                torch only owns the device memory).

`three_nn` returns SQUARED distances (as the native layer does); `knn` returns (idx (B,P,k), dist2).
"""
import numpy as np


class OracleImpl:
    name = "oracle"

    def __init__(self):
        import oracle
        self.o = oracle

    def chamfer_forward(self, x1, x2):
        return self.o.chamfer_forward(x1, x2)

    def chamfer_backward(self, x1, x2, g1, g2, i1, i2):
        return self.o.chamfer_backward(x1, x2, g1, g2, i1, i2)

    def emd_forward(self, x1, x2, eps, iters):
        return self.o.emd_forward(x1, x2, float(eps), int(iters))

    def emd_backward(self, x1, x2, g, a):
        return self.o.emd_backward(x1, x2, g, a)

    def fps(self, xyz, m):
        return self.o.furthest_point_sample(xyz, int(m))

    def fps_with_dist(self, dist, m):
        return self.o.furthest_point_sample_with_dist(dist, int(m))

    def ball_query(self, rmin, rmax, ns, xyz, centers):
        return self.o.ball_query(float(rmin), float(rmax), int(ns), xyz, centers)

    def gather(self, points, idx):
        return self.o.gather_points(points, idx)

    def gather_grad(self, go, idx, n):
        return self.o.gather_points_grad(go, idx, n)

    def group(self, points, idx):
        return self.o.group_points(points, idx)

    def group_grad(self, go, idx, n):
        return self.o.group_points_grad(go, idx, n)

    def three_nn(self, unknown, known):
        return self.o.three_nn(unknown, known)

    def three_interpolate(self, points, idx, weight):
        return self.o.three_interpolate(points, idx, weight)

    def three_interpolate_grad(self, go, idx, weight, m):
        return self.o.three_interpolate_grad(go, idx, weight, m)

    def knn(self, k, xyz, centers):
        return self.o.knn(int(k), xyz, centers)

    def knn_points(self, k, cloud, queries):
        return self.o.knn_points(int(k), cloud, queries)


class CudaImpl:
    name = "cuda"

    def __init__(self, device="cuda:0"):
        import torch
        from mvp_benchmark_b200 import _lib
        self.torch = torch
        self.L = _lib
        self.dev = torch.device(device)

    # -- helpers
    def T(self, a, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.dev)

    def E(self, shape, dtype):
        # poison outputs so that a kernel relying on pre-zeroed memory is caught
        t = self.torch.empty(shape, device=self.dev, dtype=dtype)
        if dtype == self.torch.float32:
            t.fill_(float("nan"))
        else:
            t.fill_(-12345)
        return t

    def S(self):
        import ctypes
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def N(self, *ts):
        self.torch.cuda.synchronize(self.dev)
        r = tuple(t.cpu().numpy() for t in ts)
        return r if len(r) > 1 else r[0]

    # -- ops
    ALGOS = {"auto": 0, "brute": 1, "grid": 2}

    def chamfer_forward(self, x1, x2, algo="auto"):
        t, L, p = self.torch, self.L, self.L.ptr
        a, c = self.T(x1), self.T(x2)
        b, n, _ = a.shape
        m = c.shape[1]
        d1, d2 = self.E((b, n), t.float32), self.E((b, m), t.float32)
        i1, i2 = self.E((b, n), t.int32), self.E((b, m), t.int32)
        ws = L.workspace(L.lib.mvp_chamfer_forward_workspace_bytes(b, n, m), self.dev)
        ws.fill_(0xA5)
        if algo == "auto":
            rc = L.lib.mvp_chamfer_forward(b, n, m, p(a), p(c), p(d1), p(d2), p(i1), p(i2), p(ws), ws.numel(), self.S())
        else:
            rc = L.lib.mvp_chamfer_forward_algo(self.ALGOS[algo], b, n, m, p(a), p(c), p(d1), p(d2), p(i1), p(i2), p(ws),
                                                ws.numel(), self.S())
        L.check(rc, "mvp_chamfer_forward")
        return self.N(d1, d2, i1, i2)

    BWD_ALGOS = {"auto": 0, "summed": 1}

    def chamfer_backward(self, x1, x2, g1, g2, i1, i2, algo="auto"):
        t, L, p = self.torch, self.L, self.L.ptr
        a, c = self.T(x1), self.T(x2)
        b, n, _ = a.shape
        m = c.shape[1]
        gx1, gx2 = self.E((b, n, 3), t.float32), self.E((b, m, 3), t.float32)
        tg1, tg2, ti1, ti2 = self.T(g1), self.T(g2), self.T(i1), self.T(i2)  # keep alive across the call
        L.check(L.lib.mvp_chamfer_backward_algo(self.BWD_ALGOS[algo], b, n, m, p(a), p(c), p(tg1), p(tg2), p(ti1), p(ti2),
                                                p(gx1), p(gx2), self.S()), "mvp_chamfer_backward")
        return self.N(gx1, gx2)

    EMD_ALGOS = {"auto": 0, "brute": 1, "grid": 2}

    def emd_forward(self, x1, x2, eps, iters, algo="auto"):
        t, L, p = self.torch, self.L, self.L.ptr
        a, c = self.T(x1), self.T(x2)
        b, n, _ = a.shape
        d, asg = self.E((b, n), t.float32), self.E((b, n), t.int32)
        ws = L.workspace(L.lib.mvp_emd_forward_workspace_bytes(b, n), self.dev)
        ws.fill_(0xA5)
        if algo == "auto":
            rc = L.lib.mvp_emd_forward(b, n, c.shape[1], p(a), p(c), float(eps), int(iters), p(d), p(asg), p(ws),
                                       ws.numel(), self.S())
        else:
            rc = L.lib.mvp_emd_forward_algo(self.EMD_ALGOS[algo], b, n, c.shape[1], p(a), p(c), float(eps), int(iters),
                                            p(d), p(asg), p(ws), ws.numel(), self.S())
        L.check(rc, "mvp_emd_forward")
        return self.N(d, asg)

    def emd_backward(self, x1, x2, g, asg):
        t, L, p = self.torch, self.L, self.L.ptr
        a, c = self.T(x1), self.T(x2)
        b, n, _ = a.shape
        gx = self.E((b, n, 3), t.float32)
        tg, ta = self.T(g), self.T(asg)  # keep alive across the call
        L.check(L.lib.mvp_emd_backward(b, n, p(a), p(c), p(tg), p(ta), p(gx), self.S()), "mvp_emd_backward")
        return self.N(gx)

    def fps(self, xyz, m):
        t, L, p = self.torch, self.L, self.L.ptr
        x = self.T(xyz)
        b, n, _ = x.shape
        idx = self.E((b, int(m)), t.int32)
        temp = self.E((b, n), t.float32)
        L.check(L.lib.mvp_furthest_point_sampling(b, n, int(m), p(x), p(temp), p(idx), self.S()), "mvp_fps")
        return self.N(idx)

    def fps_with_dist(self, dist, m):
        t, L, p = self.torch, self.L, self.L.ptr
        x = self.T(dist)
        b, n, _ = x.shape
        idx = self.E((b, int(m)), t.int32)
        L.check(L.lib.mvp_furthest_point_sampling_with_dist(b, n, int(m), p(x), p(None), p(idx), self.S()), "mvp_fpsd")
        return self.N(idx)

    def ball_query(self, rmin, rmax, ns, xyz, centers):
        t, L, p = self.torch, self.L, self.L.ptr
        x, c = self.T(xyz), self.T(centers)
        b, n, _ = x.shape
        m = c.shape[1]
        idx = self.E((b, m, int(ns)), t.int32)
        L.check(L.lib.mvp_ball_query(b, n, m, float(rmin), float(rmax), int(ns), p(c), p(x), p(idx), self.S()), "mvp_bq")
        return self.N(idx)

    def gather(self, points, idx):
        t, L, p = self.torch, self.L, self.L.ptr
        P, I = self.T(points), self.T(idx)
        b, c, n = P.shape
        m = I.shape[1]
        out = self.E((b, c, m), t.float32)
        L.check(L.lib.mvp_gather_points(b, c, n, m, p(P), p(I), p(out), self.S()), "mvp_gather_points")
        return self.N(out)

    def _scatter_ws(self, b, rows, entries):
        ws = self.L.workspace(self.L.lib.mvp_scatter_workspace_bytes(b, rows, entries), self.dev)
        ws.fill_(0xA5)
        return ws

    def gather_grad(self, go, idx, n, ws=False):
        """ws=True: the workspace (transposed-index) variant mvp_gather_points_grad_ws."""
        t, L, p = self.torch, self.L, self.L.ptr
        G, I = self.T(go), self.T(idx)
        b, c, m = G.shape
        out = self.E((b, c, n), t.float32)
        if ws:
            w = self._scatter_ws(b, n, m)
            L.check(L.lib.mvp_gather_points_grad_ws(b, c, n, m, p(G), p(I), p(out), p(w), w.numel(), self.S()),
                    "mvp_gather_points_grad_ws")
        else:
            L.check(L.lib.mvp_gather_points_grad(b, c, n, m, p(G), p(I), p(out), self.S()), "mvp_gather_points_grad")
        return self.N(out)

    def group(self, points, idx):
        t, L, p = self.torch, self.L, self.L.ptr
        P, I = self.T(points), self.T(idx)
        b, c, n = P.shape
        _, np_, ns = I.shape
        out = self.E((b, c, np_, ns), t.float32)
        L.check(L.lib.mvp_group_points(b, c, n, np_, ns, p(P), p(I), p(out), self.S()), "mvp_group_points")
        return self.N(out)

    def group_grad(self, go, idx, n, ws=False):
        t, L, p = self.torch, self.L, self.L.ptr
        G, I = self.T(go), self.T(idx)
        b, c, np_, ns = G.shape
        out = self.E((b, c, n), t.float32)
        if ws:
            w = self._scatter_ws(b, n, np_ * ns)
            L.check(L.lib.mvp_group_points_grad_ws(b, c, n, np_, ns, p(G), p(I), p(out), p(w), w.numel(), self.S()),
                    "mvp_group_points_grad_ws")
        else:
            L.check(L.lib.mvp_group_points_grad(b, c, n, np_, ns, p(G), p(I), p(out), self.S()), "mvp_group_points_grad")
        return self.N(out)

    def three_nn(self, unknown, known, ws=False):
        """ws=True: mvp_three_nn_ws (grid search where the shape allows it)."""
        t, L, p = self.torch, self.L, self.L.ptr
        u, k = self.T(unknown), self.T(known)
        b, n, _ = u.shape
        m = k.shape[1]
        d, idx = self.E((b, n, 3), t.float32), self.E((b, n, 3), t.int32)
        if ws:
            w = L.workspace(L.lib.mvp_three_nn_workspace_bytes(b, n, m), self.dev)
            w.fill_(0xA5)
            L.check(L.lib.mvp_three_nn_ws(b, n, m, p(u), p(k), p(d), p(idx), p(w), w.numel(), self.S()), "mvp_three_nn_ws")
        else:
            L.check(L.lib.mvp_three_nn(b, n, m, p(u), p(k), p(d), p(idx), self.S()), "mvp_three_nn")
        return self.N(d, idx)

    def three_interpolate(self, points, idx, weight):
        t, L, p = self.torch, self.L, self.L.ptr
        P, I, W = self.T(points), self.T(idx), self.T(weight)
        b, c, m = P.shape
        n = I.shape[1]
        out = self.E((b, c, n), t.float32)
        L.check(L.lib.mvp_three_interpolate(b, c, m, n, p(P), p(I), p(W), p(out), self.S()), "mvp_three_interpolate")
        return self.N(out)

    def three_interpolate_grad(self, go, idx, weight, m, ws=False):
        t, L, p = self.torch, self.L, self.L.ptr
        G, I, W = self.T(go), self.T(idx), self.T(weight)
        b, c, n = G.shape
        out = self.E((b, c, m), t.float32)
        if ws:
            w = self._scatter_ws(b, m, 3 * n)
            L.check(L.lib.mvp_three_interpolate_grad_ws(b, c, n, m, p(G), p(I), p(W), p(out), p(w), w.numel(), self.S()),
                    "mvp_three_interpolate_grad_ws")
        else:
            L.check(L.lib.mvp_three_interpolate_grad(b, c, n, m, p(G), p(I), p(W), p(out), self.S()),
                    "mvp_three_interpolate_grad")
        return self.N(out)

    def knn(self, k, xyz, centers):
        t, L, p = self.torch, self.L, self.L.ptr
        x, c = self.T(xyz), self.T(centers)
        b, n, _ = x.shape
        m = c.shape[1]
        idx, d = self.E((b, m, int(k)), t.int32), self.E((b, m, int(k)), t.float32)
        L.check(L.lib.mvp_knn(b, n, m, int(k), p(x), p(c), p(idx), p(d), self.S()), "mvp_knn")
        return self.N(idx, d)

    def knn_points(self, k, cloud, queries, grid=True):
        """grid=False: no workspace -> the exhaustive kernel."""
        t, L, p = self.torch, self.L, self.L.ptr
        c, q = self.T(cloud), self.T(queries)
        b, n, _ = q.shape
        m = c.shape[1]
        d, idx = self.E((b, n, int(k)), t.float32), self.E((b, n, int(k)), t.int32)
        if grid:
            w = L.workspace(L.lib.mvp_knn_points_workspace_bytes(b, n, m, int(k)), self.dev)
            w.fill_(0xA5)
            L.check(L.lib.mvp_knn_points(b, n, m, int(k), p(q), p(c), p(d), p(idx), p(w), w.numel(), self.S()),
                    "mvp_knn_points")
        else:
            L.check(L.lib.mvp_knn_points(b, n, m, int(k), p(q), p(c), p(d), p(idx), None, 0, self.S()), "mvp_knn_points")
        return self.N(d, idx)
