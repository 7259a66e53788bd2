function [J, idx, st, lg, stage] = bellman_sweep_multi(d, n_stages, n_gpus, opts)
%BELLMAN_SWEEP_MULTI  Backward sweep of descriptor d on n_gpus GPUs, driven by this one MATLAB process.
%   [J, idx, st, lg, stage] = bellman_sweep_multi(d, n_stages, n_gpus, opts) cuts the state grid into n_gpus slabs
%   along the dimension with the smallest halo (bellman_mex('plan', ...)), creates one handle per GPU,
%   links them (bellman_mex('group_init', ...): the stage kernels store halo values straight into the
%   neighbouring slabs' buffers over NVLink, no NCCL, no second process) and runs n_stages stages on all
%   slabs together.  Returns the value function J (S x P) and the 1-based argmin idx (S x P) of the last
%   stage computed, stitched back into the global column-major order, the run statistics, the Sigma-check
%   log (3 x checks: stage, sum(J), sum(idx); opts.check_period > 0) and the stage the run stopped at.
%   d is the struct a facade's build_desc returns; store_J_all / store_idx_all are forced off (only the
%   last stage is kept, as Solver_position / Solver_attitude / Solver_pos_att do).
    if nargin < 4, opts = struct(); end
    d.store_J_all = 0;  d.store_idx_all = 0;
    pd = bellman_mex('plan', d, n_gpus);
    hs = zeros(1, n_gpus, 'uint64');
    try
        for r = 1:n_gpus
            dr = d;  dr.part_dim = pd;  dr.rank = r - 1;  dr.nranks = n_gpus;  dr.device = r - 1;
            hs(r) = bellman_mex('create', dr);
        end
        bellman_mex('group_init', hs);
        bellman_mex('group_run', hs, n_stages, opts);
        st = bellman_mex('stats', hs(1));
        lg = bellman_mex('check_log', hs(1));
        stage = bellman_mex('current_stage', hs(1));
        n = d.n(:).';
        parts_J = cell(1, n_gpus);  parts_I = cell(1, n_gpus);
        for r = 1:n_gpus
            rg = bellman_mex('owned_range', hs(r));
            sz = n;  sz(pd) = rg(2) - rg(1);
            parts_J{r} = reshape(bellman_mex('get_J', hs(r)), [sz, d.P]);
            parts_I{r} = reshape(bellman_mex('get_idx', hs(r)), [sz, d.P]);
        end
        J = reshape(cat(pd, parts_J{:}), [], d.P);
        idx = reshape(cat(pd, parts_I{:}), [], d.P);
    catch err
        destroy_all(hs);
        rethrow(err);
    end
    destroy_all(hs);
end

function destroy_all(h)
    for k = 1:numel(h)
        if h(k) ~= 0, bellman_mex('destroy', h(k)); end
    end
end
