from ._core import lax as _lax

scan = _lax.scan
cond = _lax.cond
switch = _lax.switch
select = _lax.select
select_n = _lax.select_n
while_loop = _lax.while_loop
fori_loop = _lax.fori_loop
stop_gradient = _lax.stop_gradient
custom_linear_solve = _lax.custom_linear_solve
dynamic_slice = _lax.dynamic_slice
dynamic_update_slice = _lax.dynamic_update_slice
