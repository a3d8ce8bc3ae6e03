// capi_util.cuh -- glue shared by the C-ABI translation units.
#pragma once
#include "mab_common.cuh"
#include "asg_dev.cuh"

MabDev &mab_default_dev();                                   // lazily initialised device for the drop-in level
void mab_graph_upload(MabDev &d, const asg_t *g, DGraph &dg);
void mab_graph_download(MabDev &d, DGraph &dg, asg_t *g);
