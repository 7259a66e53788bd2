classdef Solver_pos_att < handle
    %SOLVER_POS_ATT  Coupled position+attitude channels on a 4-D (x, v, theta, w) grid, B200 back end.
    %   Drop-in for the sweep path of the reference class (pos-att/Solver_pos_att.m):
    %   simplified_run (:197-242), calculate_one_channel_U_Opt (:244-297), set_controller
    %   (:849-884), vectors_allcomb (:886-904), sym_linspace (:906-918).  fp64 throughout (the
    %   reference casts J to single); idsum50_prev starts at 0 (the reference reads it undefined).
    %   get_optimal_path (:452-500): the 13-state plant under the three channel controllers, one ode45
    %   call per stage (Dormand-Prince with MATLAB's default step control, restated), one GPU thread per X0.

    properties
        v_min
        v_max
        n_mesh_v
        x_min
        x_max
        n_mesh_x
        w_min
        w_max
        n_mesh_w
        theta1_min
        theta1_max
        theta2_min
        theta2_max
        theta3_min
        theta3_max
        n_mesh_t
        Mass
        InertiaM
        J1
        J2
        J3
        Qx1
        Qx2
        Qx3
        Qv1
        Qv2
        Qv3
        Qt1
        Qt2
        Qt3
        Qw1
        Qw2
        Qw3
        R1
        R2
        R3
        T_final
        h
        N_stage
        defaultX0
        T_dist
        F_Thr0
        F_Thr1
        F_Thr2
        F_Thr3
        F_Thr4
        F_Thr5
        F_Thr6
        F_Thr7
        F_Thr8
        F_Thr9
        F_Thr10
        F_Thr11
        Opt_F_Thr0
        Opt_F_Thr1
        Opt_F_Thr2
        Opt_F_Thr3
        Opt_F_Thr4
        Opt_F_Thr5
        Opt_F_Thr6
        Opt_F_Thr7
        Opt_F_Thr8
        Opt_F_Thr9
        Opt_F_Thr10
        Opt_F_Thr11
        device = -1
        n_gpus = 1      % > 1: every channel sweep is cut into slabs over this many GPUs, driven from this one process
    end

    methods
        function this = Solver_pos_att()
            this.v_min = -0.1;  this.v_max = 0.1;  this.n_mesh_v = 30;
            this.x_min = -0.2;  this.x_max = 0.2;  this.n_mesh_x = 30;
            this.w_min = deg2rad(-2);  this.w_max = deg2rad(2);  this.n_mesh_w = 15;
            this.theta1_min = -5; this.theta1_max = 5;
            this.theta2_min = -6; this.theta2_max = 6;
            this.theta3_min = -7; this.theta3_max = 7;
            this.n_mesh_t = 20;
            this.Mass = 4.16;
            i1 = 0.02836 + 0.00016; i2 = 0.026817 + 0.00150; i3 = 0.023 + 0.00150;
            i4 = -0.0000837; i5 = 0.000014; i6 = -0.00029;
            this.InertiaM = [i1 i4 i5; i4 i2 i6; i5 i6 i3];
            this.J1 = this.InertiaM(1); this.J2 = this.InertiaM(5); this.J3 = this.InertiaM(9);
            this.Qx1 = 6; this.Qx2 = 6; this.Qx3 = 6; this.Qv1 = 6; this.Qv2 = 6; this.Qv3 = 6;
            this.Qt1 = .5; this.Qt2 = .5; this.Qt3 = .5; this.Qw1 = .5; this.Qw2 = .5; this.Qw3 = .5;
            this.R1 = 0.1; this.R2 = 0.1; this.R3 = 0.1;
            this.T_final = 10;  this.h = 0.005;
            this.N_stage = ceil(this.T_final/this.h);
            this.defaultX0 = zeros(9,1);
            T = 0.13;  this.T_dist = 9.65E-2;
            this.F_Thr0 = [0 T];  this.F_Thr1 = [0 T];  this.F_Thr6 = -[0 T];  this.F_Thr7 = -[0 T];
            this.F_Thr2 = [0 T];  this.F_Thr3 = [0 T];  this.F_Thr8 = -[0 T];  this.F_Thr9 = -[0 T];
            this.F_Thr4 = [0 T];  this.F_Thr5 = [0 T];  this.F_Thr10 = -[0 T]; this.F_Thr11 = -[0 T];
        end

        function simplified_run(obj)
            sl = @(a,b,n) obj.sym_linspace(a, b, n);
            s_x = sl(obj.x_min, obj.x_max, obj.n_mesh_x);  s_v = sl(obj.v_min, obj.v_max, obj.n_mesh_v);
            s_w = sl(obj.w_min, obj.w_max, obj.n_mesh_w);
            s_t1 = sl(deg2rad(obj.theta1_min), deg2rad(obj.theta1_max), obj.n_mesh_t);
            s_t2 = sl(deg2rad(obj.theta2_min), deg2rad(obj.theta2_max), obj.n_mesh_t);
            s_t3 = sl(deg2rad(obj.theta3_min), deg2rad(obj.theta3_max), obj.n_mesh_t);
            obj.calculate_one_channel_U_Opt(s_x, s_v, s_t1, s_w, obj.F_Thr0, obj.F_Thr1, obj.F_Thr6, obj.F_Thr7, ...
                obj.Qx1, obj.Qv1, obj.Qt1, obj.Qw1, obj.R1, obj.J2, 'channel_x_controller_1');
            obj.calculate_one_channel_U_Opt(s_x, s_v, s_t2, s_w, obj.F_Thr2, obj.F_Thr3, obj.F_Thr8, obj.F_Thr9, ...
                obj.Qx2, obj.Qv2, obj.Qt2, obj.Qw2, obj.R2, obj.J3, 'channel_y_controller_1');
            obj.calculate_one_channel_U_Opt(s_x, s_v, s_t3, s_w, obj.F_Thr4, obj.F_Thr5, obj.F_Thr10, obj.F_Thr11, ...
                obj.Qx3, obj.Qv3, obj.Qt3, obj.Qw3, obj.R3, obj.J1, 'channel_z_controller_1');
            obj.calculate_one_channel_U_Opt(s_x, s_v, s_t1, s_w, 0, obj.F_Thr1, obj.F_Thr6, obj.F_Thr7, ...
                obj.Qx1, obj.Qv1, obj.Qt1, obj.Qw1, obj.R1, obj.J2, 'channel_x_controller_1_failure');
        end

        function calculate_one_channel_U_Opt(obj, s_x, s_v, s_t, s_w, f0, f1, f6, f7, Qx, Qv, Qt, Qw, R, J, file_name)
            [f0_allcomb, f1_allcomb, f6_allcomb, f7_allcomb] = obj.vectors_allcomb(f0, f1, f6, f7);
            s_x = s_x(:); s_v = s_v(:); s_t = s_t(:); s_w = s_w(:);  hh = obj.h;  dd = obj.T_dist;
            v_dot = (f0_allcomb + f1_allcomb + f6_allcomb + f7_allcomb)/obj.Mass;
            w_dot = (f0_allcomb*dd + f1_allcomb*(-dd) + f6_allcomb*dd + f7_allcomb*(-dd))/J;
            d.n = [numel(s_x), numel(s_v), numel(s_t), numel(s_w)];
            d.C = numel(f0_allcomb);  d.P = 1;  d.N = obj.N_stage;
            d.grid = {s_x, s_v, s_t, s_w};
            d.src_a = [1 2 3 4];  d.src_b = [2 0 4 0];
            d.Ta = {s_x, s_v, s_t, s_w};
            d.Tb = {hh*s_v, [], hh*s_w, []};
            d.Tc = {[], hh*v_dot, [], hh*w_dot};
            d.q_order = [1 2 4 3];                       % Qx x^2 + Qv v^2 + Qw w^2 + Qt t^2
            d.q  = {Qx*s_x.^2, Qv*s_v.^2, Qt*s_t.^2, Qw*s_w.^2};
            d.r  = R*f0_allcomb.^2 + R*f1_allcomb.^2 + R*f6_allcomb.^2 + R*f7_allcomb.^2;
            d.store_J_all = 0; d.store_idx_all = 0; d.device = obj.device;
            tic
            ropts = struct('check_period', 50, 'check_tol', 1e-2);
            if obj.n_gpus > 1
                [Jv, iv, ~, lg, stage_now] = bellman_sweep_multi(d, obj.N_stage - 1, obj.n_gpus, ropts);
            else
                hnd = bellman_mex('create', d);
                bellman_mex('run', hnd, obj.N_stage - 1, ropts);
                lg = bellman_mex('check_log', hnd);
                stage_now = bellman_mex('current_stage', hnd);
                Jv = bellman_mex('get_J', hnd);  iv = bellman_mex('get_idx', hnd);
                bellman_mex('destroy', hnd);
            end
            prev = [0; 0];
            for k = 1:size(lg, 2)
                fprintf('stage %d - errorF %f - errorU %f\n', lg(1,k), lg(2,k) - prev(1), lg(3,k) - prev(2));
                prev = lg(2:3,k);
            end
            if stage_now > 1
                fprintf('sum of errors in the last 50 stages is under tolerance, breaking loop...\n')
            end
            fprintf('%f seconds\n', toc)
            F_gI = griddedInterpolant({s_x.', s_v.', s_t.', s_w.'}, reshape(Jv, d.n), 'linear');
            U_Optimal_id = double(reshape(iv, d.n));
            save(file_name, 'F_gI', 'U_Optimal_id', 'f0_allcomb', 'f1_allcomb', 'f6_allcomb', 'f7_allcomb')
            fprintf('\nstage calculations complete.\n')
        end

        function set_controller(obj, file, channel)
            C = load(file);
            mk = @(f) griddedInterpolant(C.F_gI.GridVectors, f(C.U_Optimal_id), 'nearest');
            g0 = mk(C.f0_allcomb); g1 = mk(C.f1_allcomb); g6 = mk(C.f6_allcomb); g7 = mk(C.f7_allcomb);
            switch channel
                case 'x', obj.Opt_F_Thr0 = g0; obj.Opt_F_Thr1 = g1; obj.Opt_F_Thr6 = g6; obj.Opt_F_Thr7 = g7;
                case 'y', obj.Opt_F_Thr2 = g0; obj.Opt_F_Thr3 = g1; obj.Opt_F_Thr8 = g6; obj.Opt_F_Thr9 = g7;
                case 'z', obj.Opt_F_Thr4 = g0; obj.Opt_F_Thr5 = g1; obj.Opt_F_Thr10 = g6; obj.Opt_F_Thr11 = g7;
                otherwise, error('wrong channel, must be one of x-y-z values')
            end
        end

        function [X_ode45, F_Th_Opt, Force_Moment_log] = get_optimal_path(obj, X0)
            % Solver_pos_att.m:452-500 of the reference for every column of X0 (13 x batch; default the
            % reference's dr0 = [-0.1 0 0], q0 = angle2quat(0, 3 deg, 0) reversed): thruster levels from the
            % three channel controllers, moments and forces, ode45 over one stage on the 13-state plant.
            if nargin < 2
                X0 = [-0.1 0 0, 0 0 0, 0 sin(deg2rad(3)/2) 0 cos(deg2rad(3)/2), 0 0 0].';
            end
            mu = 398600;  RE = 6378;  rp = RE + 300;  e = 0.1;            % get_target_R0V0, :759-777
            ra = rp*(1 + e)/(1 - e);
            h_ = sqrt(2*mu*rp*ra/(ra + rp));
            R0 = (h_^2/mu)*(1/(1 + e))*[1 0 0];  V0 = (mu/h_)*[0 (e + 1) 0];
            files = {'channel_x_controller_1.mat', 'channel_y_controller_1.mat', 'channel_z_controller_1.mat'};   % :469-471
            hs = zeros(1, 3, 'uint64');  fv = cell(1, 3);
            for c = 1:3
                C = load(files{c});
                gv = C.F_gI.GridVectors;
                nC = numel(C.f0_allcomb);
                % a handle that only has to hold the grid and the policy: identity dynamics, zero costs
                d = struct('n', cellfun(@numel, gv), 'C', nC, 'P', 1, 'N', 2);
                d.grid = cellfun(@(g) g(:), gv, 'UniformOutput', false);
                d.src_a = [1 2 3 4];  d.src_b = [0 0 0 0];
                d.Ta = d.grid;  d.Tb = {[], [], [], []};  d.Tc = {[], zeros(nC,1), [], zeros(nC,1)};
                d.q_order = [1 2 3 4];  d.q = cellfun(@(g) zeros(numel(g),1), gv, 'UniformOutput', false);
                d.r = zeros(nC, 1);  d.store_J_all = 0;  d.store_idx_all = 0;  d.device = obj.device;
                hs(c) = bellman_mex('create', d);
                bellman_mex('set_stage', hs(c), 1, [], double(C.U_Optimal_id(:)));
                fv{c} = [C.f0_allcomb(:) C.f1_allcomb(:) C.f6_allcomb(:) C.f7_allcomb(:)];   % C_ch x 4: column m = thruster m of the channel
            end
            N = obj.N_stage;
            o = struct('n_steps', N - 1, 'stride_out', 1, 'mu', mu, 'R0', R0, 'V0', V0, 'h', obj.h, ...
                'rtol', 1e-3, 'atol', 1e-6, 'InertiaM', obj.InertiaM, 'Mass', obj.Mass, 'T_dist', obj.T_dist);
            [X, F, FM] = bellman_mex('rollout_pos_att', hs, [1 1 1], o, fv{1}, fv{2}, fv{3}, X0);
            for c = 1:3, bellman_mex('destroy', hs(c)); end
            batch = size(X0, 2);
            X_ode45 = permute(reshape(X, 13, N, batch), [2 1 3]);               % N x 13 (x batch), as :477
            F_Th_Opt = permute(reshape(F, 12, N - 1, batch), [2 1 3]);
            Force_Moment_log = permute(reshape(FM, 6, N - 1, batch), [2 1 3]);
            T_ode45 = (0:N-2)*obj.h;
            figure('Name','Thruster Firings'); plot(T_ode45, F_Th_Opt(:,:,1)); grid on
            figure('Name','states - position'); plot((0:N-1)*obj.h, X_ode45(:,1:3,1)); grid on; legend('x1','x2','x3')
            figure('Name','states - quaternions'); plot((0:N-1)*obj.h, X_ode45(:,7:10,1)); grid on; legend('q1','q2','q3','q4')
        end

        function [a1, a2, a3, a4] = vectors_allcomb(~, f1, f2, f3, f4)
            [g1, g2, g3, g4] = ndgrid(f1, f2, f3, f4);
            g1 = g1(:); g2 = g2(:); g3 = g3(:); g4 = g4(:);
            keep = ~((g1 > 0 & g3 < 0) | (g2 > 0 & g4 < 0));   % no opposing thrusters at once
            a1 = g1(keep); a2 = g2(keep); a3 = g3(keep); a4 = g4(keep);
        end

        function v = sym_linspace(~, a, b, n)
            if a > 0, error('minimum states are not negative, use normal linspace'); end
            half = ceil(n/2);
            if mod(n, 2) == 0, lo = linspace(a, 0, half + 1); else, lo = linspace(a, 0, half); end
            hi = linspace(0, b, half);
            v = [lo, hi(2:end)];
        end
    end
end
