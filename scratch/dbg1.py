import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, helpers as T
case = T.Case("cyl2d")
rows, vals = case.bc(4.0)
o, e = case.oracle(), case.engine()
o.set_dirichlet(rows, vals); e.set_dirichlet(rows); e.set_dirichlet_values(vals)
x0 = 0.3*case.random_state(); o.set_solution(x0); e.set_solution(x0)
o.assemble_first(); e.assemble_first()
Be = e.matrix("system","Bt"); Bo = T.oracle_blocks(o,"sys")["Bt"]
D = (Be - Bo).tocoo()
k = np.argsort(-np.abs(D.data))[:10]
for r,c,v in zip(D.row[k], D.col[k], D.data[k]):
    print(r, c, v, Be[r,c], Bo[r,c], r in set(rows))
print("nnz diff", (np.abs(D.data)>1e-13).sum(), "of", Be.nnz)
B2e = e.matrix("system","B"); B2o = T.oracle_blocks(o,"sys")["B"]
print("B err", T.entry_error(B2e,B2o))
