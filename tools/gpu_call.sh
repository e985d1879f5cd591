# scratch command list for gpurun calls; the round-2 measurement pass is tools/profile_r02.sh
bash tools/profile_r02.sh
