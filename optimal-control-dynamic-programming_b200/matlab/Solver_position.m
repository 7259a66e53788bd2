classdef Solver_position < handle
    %SOLVER_POSITION  Three independent (x, v) axes with three thrust levels, B200 back end.
    %   Drop-in for the sweep path of the reference class (position-control/Solver_position.m:
    %   simplified_run, :94-150).  The three axes are swept as one batched problem (P = 3) in
    %   libbellman.so; U1_Opt..U3_Opt are 'nearest' griddedInterpolants as in the reference.
    %   The orbital forward simulation (get_optimal_path, :189-361) is outside this build's scope.

    properties
        N
        Mass
        size_state_mat
        U_vector
        J_current_state_fix
        v_min
        v_max
        x_min
        x_max
        n_mesh_v
        n_mesh_x
        Qx1
        Qx2
        Qx3
        Qv1
        Qv2
        Qv3
        R1
        R2
        R3
        T_final
        h
        N_stage
        defaultX0
        U1_Opt
        U2_Opt
        U3_Opt
        F_Values        % J at the last computed stage, n_x x n_v x 3
        U_idx           % argmin (1-based), n_x x n_v x 3
        device = -1
        last_desc_ = struct()
        n_gpus = 1      % > 1: the grid is cut into slabs over this many GPUs, driven from this one process
    end

    methods
        function this = Solver_position()
            this.v_min = -0.5;  this.v_max = 0.5;
            this.x_min = -0.5;  this.x_max = 0.5;
            this.n_mesh_v = 200;  this.n_mesh_x = 200;
            this.Mass = 4.16;
            this.Qx1 = 6; this.Qx2 = 6; this.Qx3 = 6;
            this.Qv1 = 6; this.Qv2 = 6; this.Qv3 = 6;
            this.R1 = 0.1; this.R2 = 0.1; this.R3 = 0.1;
            this.T_final = 30;  this.h = 0.005;
            this.N_stage = ceil(this.T_final/this.h);
            this.defaultX0 = zeros(6,1);
            this.U_vector = [-0.13 0 0.13]*2;
        end

        function simplified_run(obj, n_stages)
            obj.N_stage = ceil(obj.T_final/obj.h);
            if nargin < 2, n_stages = obj.N_stage - 1; end
            s_x = obj.sym_linspace(obj.x_min, obj.x_max, obj.n_mesh_x).';
            s_v = obj.sym_linspace(obj.v_min, obj.v_max, obj.n_mesh_v).';
            obj.n_mesh_x = numel(s_x);  obj.n_mesh_v = numel(s_v);
            U = obj.U_vector(:);  hh = obj.h;
            % x_next = X + h*(k1+2k2+2k3+k4)/6 with k1 = V, k2 = V + k1*h/2, ... (control independent)
            k1 = s_v; k2 = s_v + k1*hh/2; k3 = s_v + k2*hh/2; k4 = s_v + k3*hh;
            incx = hh*(k1 + 2*k2 + 2*k3 + k4)/6;
            kk = U/obj.Mass;
            incv = hh*(kk + 2*kk + 2*kk + kk)/6;
            Qx = [obj.Qx1 obj.Qx2 obj.Qx3]; Qv = [obj.Qv1 obj.Qv2 obj.Qv3]; R = [obj.R1 obj.R2 obj.R3];
            rep = @(v) repmat(v, 1, 3);
            d.n = [numel(s_x), numel(s_v)];  d.C = numel(U);  d.P = 3;  d.N = obj.N_stage;
            d.grid = {rep(s_x), rep(s_v)};
            d.src_a = [1 2];  d.src_b = [2 0];  d.q_order = [1 2];
            d.Ta = {rep(s_x), rep(s_v)};
            d.Tb = {rep(incx), []};
            d.Tc = {[], rep(incv)};
            d.q  = {(s_x.^2)*Qx, (s_v.^2)*Qv};       % column a = Q_a * s.^2 (one product per element)
            d.r  = (U.^2)*R;
            d.store_J_all = 0; d.store_idx_all = 0; d.device = obj.device;
            obj.last_desc_ = d;
            sz = [numel(s_x), numel(s_v), 3];
            tic
            if obj.n_gpus > 1
                [Jv, iv] = bellman_sweep_multi(d, n_stages, obj.n_gpus, struct());
                obj.F_Values = reshape(Jv, sz);  obj.U_idx = double(reshape(iv, sz));
            else
                hnd = bellman_mex('create', d);
                bellman_mex('run', hnd, n_stages, struct('use_graph', 1));
                obj.F_Values = reshape(bellman_mex('get_J', hnd), sz);
                obj.U_idx = double(reshape(bellman_mex('get_idx', hnd), sz));
                bellman_mex('destroy', hnd);
            end
            fprintf('%d stages - %f seconds\n', n_stages, toc)
            obj.U1_Opt = griddedInterpolant({s_x.', s_v.'}, obj.U_vector(obj.U_idx(:,:,1)), 'nearest');
            obj.U2_Opt = griddedInterpolant({s_x.', s_v.'}, obj.U_vector(obj.U_idx(:,:,2)), 'nearest');
            obj.U3_Opt = griddedInterpolant({s_x.', s_v.'}, obj.U_vector(obj.U_idx(:,:,3)), 'nearest');
            fprintf('stage calculation complete!\n')
        end

        function [X_ode45, F_Opt_history] = get_optimal_path(obj, Y0)
            % Solver_position.m:189-311 of the reference: nearest policy per axis, then one rkf45 call per
            % stage on the relative-motion equations (target orbit by the universal Kepler equation), run
            % on the GPU for every column of Y0 (6 x batch; default the reference's [-1 0 0 0 0 0]').
            if nargin < 2, Y0 = [-1 0 0 0 0 0].'; end
            mu = 398600;  RE = 6378;  rp = RE + 300;  e = 0.1;            % get_target_R0V0, :313-331
            ra = rp*(1 + e)/(1 - e);
            h_ = sqrt(2*mu*rp*ra/(ra + rp));
            R0 = (h_^2/mu)*(1/(1 + e))*[1 0 0];  V0 = (mu/h_)*[0 (e + 1) 0];   % sv_from_coe at TA = RA = incl = w = 0
            N = ceil(obj.T_final/obj.h);
            d = obj.last_desc_;  d.store_J_all = 0;  d.store_idx_all = 0;
            hnd = bellman_mex('create', d);
            bellman_mex('set_stage', hnd, 1, reshape(obj.F_Values, [], 3), reshape(obj.U_idx, [], 3));
            o = struct('n_steps', N - 1, 'stride_out', 1, 'mu', mu, 'R0', R0, 'V0', V0, 'h', obj.h, 'tol', 1e-8);
            [X, id] = bellman_mex('rollout_orbit', hnd, 1, o, obj.U_vector(:), Y0);
            bellman_mex('destroy', hnd);
            batch = size(Y0, 2);
            X_ode45 = reshape(X, 6, N, batch);
            F_Opt_history = reshape(obj.U_vector(id), 3, N - 1, batch);
            T_ode45 = (0:N-1)*obj.h;
            figure; hold on; grid on; plot(T_ode45, X_ode45(1:3,:,1)); legend('x1','x2','x3')
            figure; hold on; grid on; plot(T_ode45, X_ode45(4:6,:,1)); legend('v1','v2','v3')
            figure; hold on; grid on; plot(T_ode45(1:end-1), F_Opt_history(:,:,1)); legend('u1','u2','u3')
        end

        function v = sym_linspace(~, a, b, n)
            if a > 0, error('minimum states are not negative, use normal linspace'); end
            m = ceil(n/2) + 1;
            lo = linspace(a, 0, m);  hi = linspace(0, b, m);
            v = [lo, hi(2:end)];
        end
    end
end
