import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, helpers as T
case = T.Case("cyl3d")
print("N", case.N)
o, e = case.oracle(), case.engine(precond_type="yosida")
rows, vals = case.bc(0.0)
o.set_dirichlet(rows, vals); e.set_dirichlet(rows)
x0 = case.initial(); o.set_solution(x0); e.set_solution(x0)
t=0
for step in range(3):
    t += case.dt
    rows, vals = case.bc(2.0+t)
    o.set_dirichlet_values(vals); e.set_dirichlet_values(vals)
    if step==0: o.assemble_first(); e.assemble_first()
    else: o.assemble_step(); e.assemble_step()
    print("step", step, "F err", T.entry_error(e.matrix("system","F"), T.oracle_blocks(o,"sys")["F"]), "rhs", np.abs(e.get_rhs()-o.array("rhs",case.N)).max())
    o.precond_init("yosida"); e.precond_init()
    print("  S err", T.entry_error(e.schur(), o.schur()))
    x = case.random_state(); xu, xp = x[:case.n_u], x[case.n_u:]
    print("  ilu F", T.rel_l2(e.ilu_apply(0,xu), o.ilu_apply(0,xu)), "ilu S", T.rel_l2(e.ilu_apply(1,xp), o.ilu_apply(1,xp)))
    yo = o.precond_vmult("yosida", x); ye = e.precond_vmult(x)
    print("  vmult", T.rel_l2(ye,yo), e.stat("n_inner_F"), o.stat("n_inner_F"), e.stat("n_inner_S"), o.stat("n_inner_S"))
    rc, its_o, res = o.solve_step("yosida")
    print("  oracle", rc, its_o, res, o.stat("n_inner_F"), o.stat("n_inner_S"))
    try:
        its_e,_,_ = e.solve_step()
        print("  cuda", its_e, e.stat("last_res"))
    except Exception as ex:
        print("  cuda failed", ex, e.stat("last_outer"), e.stat("last_res"), e.stat("n_inner_F"), e.stat("n_inner_S"))
        break
    print("  sol diff", T.rel_l2(e.get_solution(), o.array("sol_owned", case.N)))
