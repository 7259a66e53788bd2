classdef Dynamic_Solver < handle
    %DYNAMIC_SOLVER  Kirk ch.3 two-state linear regulator by dynamic programming, B200 back end.
    %   Drop-in for the reference class of the same name (test/Dynamic_Solver.m): same public
    %   properties, run(obj), get_optimal_path(obj, X0, mode, ssu_num), plot_u_star, compare_data.
    %   The backward sweep (Dynamic_Solver.m:86-102, 202-211 of the reference) runs in
    %   libbellman.so through bellman_mex; this file only evaluates the 1-D tables whose sums are
    %   the reference's X_next_M1/M2 and J_current_state arrays, with the same operation order.
    %
    %   Differences from the shipped reference (see DESIGN.md):
    %     * fp64 throughout (the test_coder.m / obj_1.mat convention), not single;
    %     * J_star is filled for every stage; get_optimal_path honours a supplied X0;
    %     * the S x C arrays (X1_mesh_3D, X_next_M1, J_current_state, ...) are never materialised.

    properties
        A
        B
        H
        R
        Q
        N   % number of stages
        S   % number of state values
        C   % number of control values
        x_min
        x_max
        u_min
        u_max
        s_r
        dx
        du
        u_star
        u_star_idx
        J_star
        J_current_state
        F           % struct with GridVectors and Values (J at stage 1), stands in for the interpolant
        X1_mesh
        X2_mesh
        X1_mesh_3D
        X2_mesh_3D
        X_next_M1
        X_next_M2
        U_mesh_3D
        J_check
        J_current_state_check
        J_opt_nextstate_check
        J_F_next_check
        X_next_M1_check
        X_next_M2_check
        checkstagesXJF
        % build options (not in the reference)
        U_mesh
        store_J_star = true
        device = -1
        kernel = 0
        n_gpus = 1      % > 1: slabs over this many GPUs from this one process (keeps stage 1 only: F, u_star_idx)
        handle_ = uint64(0)
    end

    methods
        function obj = Dynamic_Solver()
            obj.checkstagesXJF = 0;
            obj.Q = [0.25, 0; 0, 0.05];
            obj.A = [0.9974, 0.0539; -0.1078, 1.1591];
            obj.B = [0.0013; 0.0539];
            obj.R = 0.05;
            obj.N = 200;
            obj.S = 2;
            obj.C = 1;
            obj.dx = 100;
            obj.du = 1000;
            obj.x_max = 3;
            obj.x_min = -2.5;
            obj.u_max = 10;
            obj.u_min = -40;
        end

        function d = build_desc(obj)
            % 1-D tables; (Ta + Tb) + Tc reproduces A(1)*X1 + A(3)*X2 + B(1)*U element by element
            s = linspace(obj.x_min, obj.x_max, obj.dx).';
            u = linspace(obj.u_min, obj.u_max, obj.du).';
            obj.s_r = s.';
            obj.U_mesh = u.';
            d.n = [obj.dx, obj.dx];
            d.C = obj.du;  d.P = 1;  d.N = obj.N;
            d.grid = {s, s};
            d.src_a = [1 1];  d.src_b = [2 2];  d.q_order = [1 2];
            d.Ta = {obj.A(1)*s, obj.A(2)*s};
            d.Tb = {obj.A(3)*s, obj.A(4)*s};
            d.Tc = {obj.B(1)*u, obj.B(2)*u};
            d.q  = {obj.Q(1)*s.^2, obj.Q(4)*s.^2};
            d.r  = obj.R*u.^2;
            d.store_J_all = obj.store_J_star;  d.store_idx_all = 1;
            d.idx_bytes = 1 + (obj.du > 256) + 2*(obj.du > 65536);   % u_star of every stage stays on the device: 1 or 2 bytes per state
            d.device = obj.device;
        end

        function obj = run(obj)
            if obj.handle_ ~= 0, bellman_mex('destroy', obj.handle_); obj.handle_ = uint64(0); end
            d = obj.build_desc();
            [obj.X1_mesh, obj.X2_mesh] = ndgrid(obj.s_r, obj.s_r);
            if obj.n_gpus > 1
                tic
                [Jv, iv, st] = bellman_sweep_multi(d, obj.N - 1, obj.n_gpus, struct('kernel', obj.kernel));
                fprintf('%d stages on %d GPUs - %f seconds (device %.3f ms, %s kernel)\n', obj.N - 1, obj.n_gpus, toc, st.ms, st.kernel)
                obj.u_star_idx = double(reshape(iv, [obj.dx, obj.dx]));
                obj.F = struct('GridVectors', {{obj.s_r, obj.s_r}}, 'Values', reshape(Jv, [obj.dx, obj.dx]));
                return
            end
            obj.handle_ = bellman_mex('create', d);
            tic
            bellman_mex('run', obj.handle_, obj.N - 1, struct('kernel', obj.kernel));
            st = bellman_mex('stats', obj.handle_);
            fprintf('%d stages - %f seconds (device %.3f ms, %s kernel)\n', obj.N - 1, toc, st.ms, st.kernel)
            sz = [obj.dx, obj.dx];
            if obj.store_J_star
                obj.J_star = zeros([sz, obj.N]);
                obj.u_star = zeros([sz, obj.N]);
                for k = 1:obj.N
                    obj.J_star(:,:,k) = reshape(bellman_mex('get_J', obj.handle_, k), sz);
                end
                for k = 1:obj.N-1
                    obj.u_star(:,:,k) = obj.U_mesh(reshape(bellman_mex('get_idx', obj.handle_, k), sz));
                end
            end
            obj.u_star_idx = double(reshape(bellman_mex('get_idx', obj.handle_, 1), sz));
            obj.F = struct('GridVectors', {{obj.s_r, obj.s_r}}, ...
                           'Values', reshape(bellman_mex('get_J', obj.handle_, 1), sz));
        end

        function [X, U] = get_optimal_path(obj, X0, mode, ssu_num)
            if nargin < 2 || isempty(X0), X0 = [2; 1]; end
            if nargin < 3, mode = 'Nssu'; ssu_num = 1; end
            if nargin < 4, ssu_num = 1; end
            X0 = reshape(X0, 2, []);
            [Xf, U] = bellman_mex('rollout', obj.handle_, obj.N, obj.A, obj.B, obj.U_mesh, X0, ...
                                  double(strcmp(mode, 'ssu')), ssu_num);
            X = reshape(Xf, 2, obj.N, []);
            if size(X0, 2) == 1 && nargout == 0
                v = 1:obj.N;
                plot(v, X(1,v)); hold on
                plot(v, X(2,v), 'r'); plot(v, U(v), '--')
                title('Optimal control for initial state X0')
                xlabel('stage - k'); ylabel('state and inputs')
                legend('X1', 'X2', 'u*'); grid on; xlim([v(1) v(end)])
            end
        end

        function plot_u_star(this, k_s)
            if nargin < 2, k_s = 1:this.N-2; end
            figure
            if numel(k_s) == 1
                plot3(this.X1_mesh, this.X2_mesh, this.u_star(:,:,k_s))
            else
                p = mesh(this.X1_mesh, this.X2_mesh, this.u_star(:,:,k_s(1)));
                colormap winter; axis manual
                for i = 2:numel(k_s)
                    p.ZData = this.u_star(:,:,k_s(i));
                    title(['Stage ', num2str(k_s(i))]); pause(0.2)
                end
            end
        end

        function delete(obj)
            if obj.handle_ ~= 0, bellman_mex('destroy', obj.handle_); end
        end
    end

    methods (Static)
        function b = compare_data(obj1, obj2)
            if isempty(obj1.J_star) || isempty(obj2.J_star)
                error('stop throwing empty data at me')
            end
            b = isequal(obj1.J_star, obj2.J_star);
            if b, disp('J_star matrices comparison -- Match!')
            else, warning('J_star matrices -- Do NOT match'); end
        end
    end
end
