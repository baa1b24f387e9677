// K2/K3 statistics kernels -- see abi_stats.inl.  (filled in next)
#pragma once
