"""Development helper: columns of the diagonal-format operator (apply it to I + e_j) against the dense float64 oracle."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import matting
pkg = "automated-deep-photo-style-transfer_b200"
v2 = importlib.import_module(pkg + ".components.matting_v2"); synth = importlib.import_module(pkg + ".synth")
H, W = 6, 9
for kind in ("smooth", "grey"):
    img32 = synth.smooth_image(8, 16, 48)[0][:H, :W]
    if kind == "grey":
        img32 = np.repeat(img32[..., :1], 3, -1)
    img32 = np.ascontiguousarray(img32)
    ref = matting.V2Operator(img32.astype(np.float64), 1e-7, 1)
    A = np.asarray(ref.matmul(np.eye(H * W)))
    LIo = ref.matmul(img32.reshape(-1, 3).astype(np.float64))
    for kern in ("dia", "matrix_free"):
        op = v2.MattingLaplacian(torch.as_tensor(img32).cuda(), epsilon=1e-7, storage_dtype=torch.float32, compute_dtype=torch.float64, kernel=kern)
        I = torch.as_tensor(img32.reshape(-1, 3)).cuda()
        yI, _ = op.quadratic_form(I, want_y=True)
        yI = yI.cpu().double().numpy()
        print(kind, kern, "LI err %.2e of %.2e" % (np.abs(yI - LIo).max(), np.abs(LIo).max()))
        Ag = np.zeros((H * W, H * W))
        for j in range(H * W):
            x = I.clone(); x[j, 1] += 0.25
            y, _ = op.quadratic_form(x, want_y=True)
            Ag[:, j] = (y.cpu().double().numpy()[:, 1] - yI[:, 1]) / 0.25
        E = np.abs(Ag - A)
        i, j = np.unravel_index(E.argmax(), E.shape)
        print(kind, kern, "column err max %.2e at i=(%d,%d) j=(%d,%d) A=%.6f got=%.6f; offdiag max %.2e diag max %.2e" % (
            E.max(), i // W, i % W, j // W, j % W, A[i, j], Ag[i, j], (E - np.diag(np.diag(E))).max(), np.diag(E).max()))
        x = torch.rand(H * W, 3, device="cuda")
        y, _ = op.quadratic_form(x, want_y=True)
        want = ref.matmul(x.cpu().double().numpy())
        print(kind, kern, "random x err %.2e of %.2e" % (np.abs(y.cpu().double().numpy() - want).max(), np.abs(want).max()))
