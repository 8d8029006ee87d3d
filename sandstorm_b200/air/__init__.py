"""Host side of constraint evaluation: the expression vocabulary of ministark's
`Expr<AlgebraicItem<..>>` as the reference's AirConfig::constraints uses it
(layouts/src/recursive/air.rs:61-81: X, Constant, Trace(col, offset), Challenge, Hint, Periodic;
+ - * / pow), the composition Σ constraint_i * alpha^i (air.rs:1184-1200), and the compiler that
flattens the DAG into the straight-line program `ss_constraint_eval` executes on the GPU."""
from .expr import (Challenge, Constant, Expr, Hint, Periodic, Trace, X, composition_constraint)
from .program import CompiledProgram, ProgramTemplate, compile_program, compile_template

__all__ = ["Expr", "X", "Constant", "Trace", "Challenge", "Hint", "Periodic", "composition_constraint",
           "compile_program", "CompiledProgram", "compile_template", "ProgramTemplate"]
